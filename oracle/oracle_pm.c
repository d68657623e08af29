/* oracle_pm.c -- CPU restatement of the MP-Gadget particle-mesh arithmetic.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The FFTs themselves are done by the Python harness with scipy.fft
 * (pocketfft), unnormalised in both directions as PFFT/FFTW are
 * (petapm.c:289-293,1169-1182).  The third-party FFT of the reference is
 * PFFT 1.0.8-alpha3-fftw3-2don2d over FFTW3 (depends/install_pfft.sh:9),
 * which is not available offline; any correct DFT is equivalent up to rounding.
 *
 * Arrays here are on the GLOBAL mesh (x slowest, z fastest); the reference's
 * per-region buffers + pencil exchange (petapm.c:584-930) only redistribute the
 * same cell sums.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "oracle.h"

static inline int wrap(int i, int N)       /* petapm.c:903-906,916-918 */
{
    while(i < 0) i += N;
    while(i >= N) i -= N;
    return i;
}

/* pm_iterate_one petapm.c:955-1006: iCell = floor(Pos/CellSize), Res = frac,
 * 8 connections with weight prod_k (offset ? Res : 1-Res), k = 0,1,2. */
static inline void cic_stencil(const double *Pos, double CellSize, int iCell[3], double Res[3])
{
    for(int k = 0; k < 3; k++) {
        double tmp = Pos[k] / CellSize;
        iCell[k] = (int) floor(tmp);
        Res[k] = tmp - iCell[k];
    }
}

void oracle_pm_deposit(const double *pos, const float *mass, int64_t n, double BoxSize, int Nmesh,
                       double *mesh, int32_t *icell_out)
{
    const double CellSize = BoxSize / Nmesh;            /* petapm.c:112 */
#pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        int iCell[3]; double Res[3];
        cic_stencil(&pos[3 * i], CellSize, iCell, Res);
        if(icell_out)
            for(int k = 0; k < 3; k++) icell_out[3 * i + k] = iCell[k];
        const double Mass = (double) mass[i];            /* petapm.c:1139 */
        for(int c = 0; c < 8; c++) {
            double weight = 1.0;
            int64_t linear = 0;
            for(int k = 0; k < 3; k++) {
                int offset = (c >> k) & 1;
                int ix = wrap(iCell[k] + offset, Nmesh);
                linear = linear * Nmesh + ix;
                weight *= offset ? Res[k] : (1 - Res[k]);
            }
            const double add = weight * Mass;
#pragma omp atomic update
            mesh[linear] += add;                          /* petapm.c:1142-1143 */
        }
    }
}

void oracle_pm_readout(const double *mesh, const double *pos, int64_t n, double BoxSize, int Nmesh,
                       double *out, int64_t ostride)
{
    const double CellSize = BoxSize / Nmesh;
#pragma omp parallel for
    for(int64_t i = 0; i < n; i++) {
        int iCell[3]; double Res[3];
        cic_stencil(&pos[3 * i], CellSize, iCell, Res);
        for(int c = 0; c < 8; c++) {
            double weight = 1.0;
            int64_t linear = 0;
            for(int k = 0; k < 3; k++) {
                int offset = (c >> k) & 1;
                int ix = wrap(iCell[k] + offset, Nmesh);
                linear = linear * Nmesh + ix;
                weight *= offset ? Res[k] : (1 - Res[k]);
            }
            out[i * ostride] += weight * mesh[linear];    /* gravpm.c:499-510 */
        }
    }
}

static double sinc_unnormed(double x)                    /* gravpm.c:295-302 */
{
    if(x < 1e-5 && x > -1e-5) {
        double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

static inline int mesh_to_k(int i, int N) { return i <= N / 2 ? i : i - N; }   /* petapm.c:81-84 */

void oracle_pm_potential_transfer(double *rhok, int Nmesh, double BoxSize, double Asmth, double G)
{
    const int Nz = Nmesh / 2 + 1;
    const double asmth2 = pow((2 * M_PI) * Asmth / Nmesh, 2);          /* gravpm.c:386 */
    const double pot_factor = -G / (M_PI * BoxSize);                    /* gravpm.c:392 */
#pragma omp parallel for collapse(2)
    for(int ix = 0; ix < Nmesh; ix++)
        for(int iy = 0; iy < Nmesh; iy++)
            for(int iz = 0; iz < Nz; iz++) {
                const int kpos[3] = {mesh_to_k(ix, Nmesh), mesh_to_k(iy, Nmesh), mesh_to_k(iz, Nmesh)};
                int64_t k2 = 0;
                for(int k = 0; k < 3; k++) k2 += ((int64_t) kpos[k]) * kpos[k];
                double *v = &rhok[2 * (((int64_t) ix * Nmesh + iy) * Nz + iz)];
                if(k2 == 0) { v[0] = 0.0; v[1] = 0.0; continue; }       /* gravpm.c:441-449 */
                double f = 1.0;
                const double smth = exp(-k2 * asmth2) / k2;             /* gravpm.c:388 */
                for(int k = 0; k < 3; k++) {                             /* gravpm.c:403-407 */
                    double tmp = (kpos[k] * M_PI) / Nmesh;
                    tmp = sinc_unnormed(tmp);
                    f *= 1. / (tmp * tmp);
                }
                const double fac = pot_factor * smth * f * f;           /* gravpm.c:413 */
                v[0] *= fac;
                v[1] *= fac;
            }
}

/* powerspectrum_add_mode over all modes of the untouched density spectrum
 * (gravpm.c:330-361, called from potential_transfer gravpm.c:440): raw sums of the
 * nbins = Nmesh logarithmic bins, before powerspectrum_sum (powerspectrum.c:56-92). */
void oracle_pm_power(const double *rhok, int Nmesh, double *power, double *kk, int64_t *nmodes, double *norm)
{
    const int Nz = Nmesh / 2 + 1, size = Nmesh;
    for(int b = 0; b < size; b++) { power[b] = 0; kk[b] = 0; nmodes[b] = 0; }
    const double binsperunit = (size - 1) / log(sqrt(3) * Nmesh / 2.0);
    for(int ix = 0; ix < Nmesh; ix++)
        for(int iy = 0; iy < Nmesh; iy++)
            for(int iz = 0; iz < Nz; iz++) {
                const int kpos[3] = {mesh_to_k(ix, Nmesh), mesh_to_k(iy, Nmesh), mesh_to_k(iz, Nmesh)};
                int64_t k2 = 0;
                for(int k = 0; k < 3; k++) k2 += ((int64_t) kpos[k]) * kpos[k];
                const double *v = &rhok[2 * (((int64_t) ix * Nmesh + iy) * Nz + iz)];
                const double m = v[0] * v[0] + v[1] * v[1];
                if(k2 == 0) { *norm = m; continue; }
                double f = 1.0;
                for(int k = 0; k < 3; k++) {
                    double tmp = (kpos[k] * M_PI) / Nmesh;
                    tmp = sinc_unnormed(tmp);
                    f *= 1. / (tmp * tmp);
                }
                const int kint = floor(binsperunit * log(k2) / 2.);
                if(kint >= size) continue;
                const int w = (kpos[2] == 0 || kpos[2] == Nmesh / 2) ? 1 : 2;
                const double keff = sqrt(kpos[0] * kpos[0] + kpos[1] * kpos[1] + kpos[2] * kpos[2]);
                power[kint] += w * m * f * f;
                nmodes[kint] += w;
                kk[kint] += w * keff;
            }
}

static double diff_kernel(double w)                       /* gravpm.c:458-466 */
{
    return 1 / 6.0 * (8 * sin(w) - sin(2 * w));
}

void oracle_pm_force_transfer(const double *potk, double *out, int Nmesh, double BoxSize, int dim)
{
    const int Nz = Nmesh / 2 + 1;
#pragma omp parallel for collapse(2)
    for(int ix = 0; ix < Nmesh; ix++)
        for(int iy = 0; iy < Nmesh; iy++)
            for(int iz = 0; iz < Nz; iz++) {
                const int kpos[3] = {mesh_to_k(ix, Nmesh), mesh_to_k(iy, Nmesh), mesh_to_k(iz, Nmesh)};
                const int64_t o = 2 * (((int64_t) ix * Nmesh + iy) * Nz + iz);
                /* force_transfer gravpm.c:476-489 */
                const double fac = -1 * diff_kernel(kpos[dim] * (2 * M_PI / Nmesh)) * (Nmesh / BoxSize);
                const double re = potk[o], im = potk[o + 1];
                out[o] = -im * fac;
                out[o + 1] = re * fac;
            }
}
