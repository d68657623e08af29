/* pfft_standin.c -- the PFFT entry points libgadget/petapm.c calls, for ONE process, as plain
 * separable discrete Fourier transforms.  TEST INFRASTRUCTURE ONLY: it lets the reference's own
 * petapm.c / gravpm.c run unmodified (oracle/Makefile.ref, target libref_pm.so) so that the PM
 * oracle and the CUDA PM path can be pinned against the reference's mesh arithmetic.  The FFT is
 * the one part of that pipeline with a unique mathematical answer, so replacing the library by a
 * direct evaluation changes nothing but rounding (~1e-15 relative).
 *
 * Conventions of PFFT/FFTW reproduced here (petapm.c:147-187,289-293,1169-1182):
 *   forward  r2c: F[k] = sum_x f[x] exp(-2 pi i k.x / N), unnormalised;
 *   backward c2r: f[x] = sum_k F[k] exp(+2 pi i k.x / N), unnormalised, Hermitian completion in z;
 *   real array [x][y][z] contiguous (local_ni = {N,N,N});
 *   PFFT_TRANSPOSED_OUT / _IN: the half spectrum is stored in the order (y, z, x), z in [0, N/2],
 *   and pfft_local_size reports local_no in (x, y, z) order = {N, N, N/2+1} (petapm.c rolls it). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "stubs/pfft.h"

struct pfft_plan_s { int N; int r2c; };

void pfft_init(void) {}
void pfft_plan_with_nthreads(int nthreads) { (void) nthreads; }
int pfft_create_procmesh_2d(MPI_Comm comm, int np0, int np1, MPI_Comm *cart)
{
    if(np0 * np1 != 1) return 1;
    *cart = comm;
    return 0;
}

ptrdiff_t pfft_local_size_dft_r2c_3d(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned flags,
                                     ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no, ptrdiff_t *local_o_start)
{
    (void) comm_cart; (void) flags;
    for(int d = 0; d < 3; d++) { local_ni[d] = n[d]; local_i_start[d] = 0; local_no[d] = n[d]; local_o_start[d] = 0; }
    local_no[2] = n[2] / 2 + 1;
    return n[0] * n[1] * (n[2] / 2 + 1);
}

static pfft_plan make_plan(const ptrdiff_t *n, int r2c)
{
    if(n[0] != n[1] || n[1] != n[2]) return NULL;
    pfft_plan p = (pfft_plan) malloc(sizeof(*p));
    p->N = (int) n[0]; p->r2c = r2c;
    return p;
}
pfft_plan pfft_plan_dft_r2c_3d(const ptrdiff_t *n, double *in, pfft_complex *out, MPI_Comm c, int sign, unsigned flags)
{
    (void) in; (void) out; (void) c; (void) sign; (void) flags;
    return make_plan(n, 1);
}
pfft_plan pfft_plan_dft_c2r_3d(const ptrdiff_t *n, pfft_complex *in, double *out, MPI_Comm c, int sign, unsigned flags)
{
    (void) in; (void) out; (void) c; (void) sign; (void) flags;
    return make_plan(n, 0);
}
void pfft_destroy_plan(pfft_plan plan) { free(plan); }

/* twiddles w[j] = exp(sign * 2 pi i j / N), j in [0, N) */
static void twiddles(int N, int sign, double *wr, double *wi)
{
    for(int j = 0; j < N; j++) {
        const double a = 2.0 * M_PI * (double) j / (double) N;
        wr[j] = cos(a); wi[j] = sign * sin(a);
    }
}

/* one pass: out[k] = sum_j in[j] w^(j k) along an axis with the given stride, for `lines` lines */
static void dft_axis(int N, int nk, const double *inr, const double *ini, double *outr, double *outi,
                     size_t nlines, const size_t *line_in, const size_t *line_out, size_t stride_in, size_t stride_out,
                     const double *wr, const double *wi)
{
#pragma omp parallel for schedule(static)
    for(size_t l = 0; l < nlines; l++) {
        const double *ar = inr + line_in[l], *ai = ini ? ini + line_in[l] : NULL;
        for(int k = 0; k < nk; k++) {
            double sr = 0, si = 0;
            for(int j = 0; j < N; j++) {
                const int t = (int) (((long long) j * k) % N);
                const double xr = ar[j * stride_in], xi = ai ? ai[j * stride_in] : 0.0;
                sr += xr * wr[t] - xi * wi[t];
                si += xr * wi[t] + xi * wr[t];
            }
            outr[line_out[l] + k * stride_out] = sr;
            outi[line_out[l] + k * stride_out] = si;
        }
    }
}

void pfft_execute_dft_r2c(const pfft_plan plan, double *in, pfft_complex *out)
{
    const int N = plan->N, Nz = N / 2 + 1;
    const size_t n3 = (size_t) N * N * N, nh = (size_t) N * N * Nz;
    double *wr = (double *) malloc(sizeof(double) * 2 * N), *wi = wr + N;
    twiddles(N, -1, wr, wi);
    /* work arrays in [x][y][kz] order */
    double *ar = (double *) malloc(sizeof(double) * 4 * nh), *ai = ar + nh, *br = ai + nh, *bi = br + nh;
    size_t *li = (size_t *) malloc(sizeof(size_t) * 2 * (size_t) N * N), *lo = li + (size_t) N * N;
    /* z: real -> half complex */
    for(size_t q = 0; q < (size_t) N * N; q++) { li[q] = q * N; lo[q] = q * Nz; }
    dft_axis(N, Nz, in, NULL, ar, ai, (size_t) N * N, li, lo, 1, 1, wr, wi);
    /* y */
    size_t *li2 = (size_t *) malloc(sizeof(size_t) * (size_t) N * Nz);
    for(int x = 0; x < N; x++) for(int z = 0; z < Nz; z++) li2[(size_t) x * Nz + z] = (size_t) x * N * Nz + z;
    dft_axis(N, N, ar, ai, br, bi, (size_t) N * Nz, li2, li2, Nz, Nz, wr, wi);
    /* x, written transposed: out[y][z][x] */
    size_t *li3 = (size_t *) malloc(sizeof(size_t) * 2 * (size_t) N * Nz), *lo3 = li3 + (size_t) N * Nz;
    for(int y = 0; y < N; y++) for(int z = 0; z < Nz; z++) {
        li3[(size_t) y * Nz + z] = (size_t) y * Nz + z;              /* [x][y][z] with x stride N*Nz */
        lo3[(size_t) y * Nz + z] = ((size_t) y * Nz + z) * N;        /* [y][z][x] with x stride 1 */
    }
    dft_axis(N, N, br, bi, ar, ai, (size_t) N * Nz, li3, lo3, (size_t) N * Nz, 1, wr, wi);
    for(size_t q = 0; q < nh; q++) { out[q][0] = ar[q]; out[q][1] = ai[q]; }
    (void) n3;
    free(li3); free(li2); free(li); free(ar); free(wr);
}

void pfft_execute_dft_c2r(const pfft_plan plan, pfft_complex *in, double *out)
{
    const int N = plan->N, Nz = N / 2 + 1;
    const size_t nh = (size_t) N * N * Nz;
    double *wr = (double *) malloc(sizeof(double) * 2 * N), *wi = wr + N;
    twiddles(N, +1, wr, wi);
    double *ar = (double *) malloc(sizeof(double) * 4 * nh), *ai = ar + nh, *br = ai + nh, *bi = br + nh;
    for(size_t q = 0; q < nh; q++) { ar[q] = in[q][0]; ai[q] = in[q][1]; }       /* [y][z][x] */
    /* x: [y][z][kx] -> [x][y][z]-ordered work array (half spectrum still in z) */
    size_t *li = (size_t *) malloc(sizeof(size_t) * 2 * (size_t) N * Nz), *lo = li + (size_t) N * Nz;
    for(int y = 0; y < N; y++) for(int z = 0; z < Nz; z++) {
        li[(size_t) y * Nz + z] = ((size_t) y * Nz + z) * N;
        lo[(size_t) y * Nz + z] = (size_t) y * Nz + z;
    }
    dft_axis(N, N, ar, ai, br, bi, (size_t) N * Nz, li, lo, 1, (size_t) N * Nz, wr, wi);
    /* y */
    for(int x = 0; x < N; x++) for(int z = 0; z < Nz; z++) li[(size_t) x * Nz + z] = (size_t) x * N * Nz + z;
    dft_axis(N, N, br, bi, ar, ai, (size_t) N * Nz, li, li, Nz, Nz, wr, wi);
    /* z: Hermitian half spectrum -> real.  f[z] = sum_{kz=0}^{N-1} F[kz] w^(kz z), F[N-kz] = conj(F[kz]) */
#pragma omp parallel for schedule(static)
    for(size_t q = 0; q < (size_t) N * N; q++) {
        const double *fr = ar + q * Nz, *fi = ai + q * Nz;
        for(int z = 0; z < N; z++) {
            double s = 0;
            for(int k = 0; k < N; k++) {
                const int kk = k < Nz ? k : N - k;
                const double xr = fr[kk], xi = k < Nz ? fi[kk] : -fi[kk];
                const int t = (int) (((long long) k * z) % N);
                s += xr * wr[t] - xi * wi[t];
            }
            out[q * N + z] = s;
        }
    }
    free(li); free(ar); free(wr);
}
