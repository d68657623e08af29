// pm.cu -- long-range particle-mesh gravity on one B200.
//
// Replaces gravpm_force / petapm_force (libgadget/gravpm.c:60-119,
// petapm.c:263-379) for gravpm's callback set.  The reference deposits into
// per-region buffers and ships compressed pencils to a 2-D PFFT decomposition
// (petapm.c:584-930); on one GPU the whole Nmesh^3 mesh is resident in HBM, so
// particles deposit straight into the global mesh.
//
// Pipeline (all on E->stream):
//   clear mesh -> CIC deposit (fp64 RED atomics)              petapm.c:955-1006,1138-1144
//   cuFFT D2Z                                                   petapm.c:305
//   Green's function in place on the half spectrum             gravpm.c:383-454
//   cuFFT Z2D -> potential mesh                                 petapm.c:344
//   4-point finite-difference gradient -> 3 force meshes
//   one trilinear gather of the 4 fields per particle           gravpm.c:499-510
//
// The reference obtains the force meshes by multiplying the potential spectrum
// by i*D(w), D(w) = (8 sin w - sin 2w)/6 * Nmesh/Box (gravpm.c:458-489) and
// running three more inverse FFTs.  i*D(w)/h is exactly the transfer function
// of the 4-point central difference [8(f(x+h)-f(x-h)) - (f(x+2h)-f(x-2h))]/(12h)
// ("the same as GADGET-2 but in fourier space", gravpm.c:460-464), so taking
// that difference of the potential mesh in real space gives the same force
// meshes to rounding with 2 FFTs per step instead of 5.
#include "engine.h"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

namespace b200 {

static const char *cufft_str(cufftResult r)
{
    switch(r) {
        case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
        case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
        case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
        case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
        case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
        case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
        case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
        case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
        default: return "CUFFT_ERROR";
    }
}
#define CKF(call) do { cufftResult _r = (call); if(_r != CUFFT_SUCCESS) return failmsg(E, std::string(#call) + ": " + cufft_str(_r)); } while(0)

__device__ __forceinline__ int wrapi(int i, int N)
{
    i %= N;
    return i < 0 ? i + N : i;
}

// CIC stencil of pm_iterate_one (petapm.c:976-980): tmp = Pos/CellSize,
// iCell = floor(tmp), Res = tmp - iCell.  The division is kept (not replaced by
// a reciprocal) so that iCell is bit-identical to the reference.
__device__ __forceinline__ void cic_cell(const double *__restrict__ p, double cellsize, int ic[3], double res[3])
{
#pragma unroll
    for(int k = 0; k < 3; k++) {
        const double t = __ddiv_rn(p[k], cellsize);
        const double f = floor(t);
        ic[k] = (int) f;
        res[k] = t - f;
    }
}

__global__ void __launch_bounds__(256)
k_pm_deposit(const double *__restrict__ pos, const float *__restrict__ mass,
             const uint8_t *__restrict__ flags, int64_t n, double cellsize, int N,
             double *__restrict__ mesh)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(flags[i] & 3) return;      // swallowed / garbage: RegionInd = -2 (gravpm.c:176-178)
    int ic[3]; double res[3];
    cic_cell(pos + 3 * i, cellsize, ic, res);
    const double m = (double) mass[i];
    const int x0 = wrapi(ic[0], N), x1 = wrapi(ic[0] + 1, N);
    const int y0 = wrapi(ic[1], N), y1 = wrapi(ic[1] + 1, N);
    const int z0 = wrapi(ic[2], N), z1 = wrapi(ic[2] + 1, N);
    const double wx[2] = {1 - res[0], res[0]};
    const double wy[2] = {1 - res[1], res[1]};
    const double wz[2] = {1 - res[2], res[2]};
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1}, zs[2] = {z0, z1};
#pragma unroll
    for(int c = 0; c < 8; c++) {
        const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
        // weight = ((1*wx)*wy)*wz, then * Mass  (petapm.c:991-1000,1143)
        const double w = __dmul_rn(__dmul_rn(__dmul_rn(wx[ox], wy[oy]), wz[oz]), m);
        const size_t lin = ((size_t) xs[ox] * N + ys[oy]) * N + zs[oz];
        atomicAdd(mesh + lin, w);     // RED.E.ADD.F64
    }
}

__global__ void __launch_bounds__(256)
k_pm_cell_index(const double *__restrict__ pos, int64_t n, double cellsize, int32_t *__restrict__ out)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    int ic[3]; double res[3];
    cic_cell(pos + 3 * i, cellsize, ic, res);
    out[3 * i] = ic[0]; out[3 * i + 1] = ic[1]; out[3 * i + 2] = ic[2];
}

// potential_transfer (gravpm.c:383-454) applied in place to the half spectrum,
// layout [ix][iy][iz], iz in [0, N/2].  ktab[i] = 1/sinc^2(pi k_i / N).
//
// POWER: also powerspectrum_add_mode (gravpm.c:330-361) on the untouched density modes: bin
// floor(binsperunit * log(k2) / 2) of N bins, weight 2 except in the planes kz = 0 and N/2,
// value |rho_k|^2 deconvolved by f^2; the zero mode is the normalisation.  Bins are summed per
// block in shared memory, then once per block into ps[3][N] = {Power, kk, Nmodes}, ps[3N] = Norm.
template <bool POWER>
__global__ void __launch_bounds__(256)
k_pm_potential_transfer(double2 *__restrict__ v, int N, int Nz, const double *__restrict__ ktab,
                        double asmth2, double pot_factor, double binsperunit, double *__restrict__ ps)
{
    extern __shared__ double s_ps[];            // POWER: [3][N]
    if(POWER) {
        for(int b = threadIdx.x; b < 3 * N; b += blockDim.x) s_ps[b] = 0;
        __syncthreads();
    }
    const size_t total = (size_t) N * N * Nz;
    for(size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total;
        idx += (size_t) gridDim.x * blockDim.x) {
        const int iz = (int) (idx % Nz);
        const size_t row = idx / Nz;
        const int iy = (int) (row % N);
        const int ix = (int) (row / N);
        const int kx = ix <= N / 2 ? ix : ix - N;       // petapm_mesh_to_k petapm.c:81-84
        const int ky = iy <= N / 2 ? iy : iy - N;
        const int kz = iz;
        const long long k2 = (long long) kx * kx + (long long) ky * ky + (long long) kz * kz;
        double2 val = v[idx];
        if(k2 == 0) {
            if(POWER) ps[3 * N] = val.x * val.x + val.y * val.y;       // gravpm.c:332-336
            val.x = 0.0; val.y = 0.0;                   // gravpm.c:441-449
        } else {
            const double smth = exp((double) (-k2) * asmth2) / (double) k2;
            const double f = (ktab[ix] * ktab[iy]) * ktab[iz];
            if(POWER) {
                const int kint = (int) floor(binsperunit * log((double) k2) / 2.);
                if(kint < N) {
                    const double w = (kz == 0 || kz == N / 2) ? 1.0 : 2.0;
                    const double m = val.x * val.x + val.y * val.y;
                    atomicAdd(&s_ps[kint], w * m * f * f);
                    atomicAdd(&s_ps[N + kint], w * sqrt((double) k2));
                    atomicAdd(&s_ps[2 * N + kint], w);
                }
            }
            const double fac = ((pot_factor * smth) * f) * f;
            val.x *= fac; val.y *= fac;
        }
        v[idx] = val;
    }
    if(POWER) {
        __syncthreads();
        for(int b = threadIdx.x; b < 3 * N; b += blockDim.x) if(s_ps[b] != 0) atomicAdd(&ps[b], s_ps[b]);
    }
}

// F_d = -[8(P(x+h)-P(x-h)) - (P(x+2h)-P(x-2h))]/(12h), see file header.
__global__ void __launch_bounds__(256)
k_pm_gradient(const double *__restrict__ pot, int N, double inv12h,
              double *__restrict__ fx, double *__restrict__ fy, double *__restrict__ fz)
{
    const int iz = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y;
    const int ix = blockIdx.z;
    if(iz >= N) return;
    const size_t NN = (size_t) N * N;
    const size_t base = (size_t) ix * NN + (size_t) iy * N + iz;
    const int xm1 = wrapi(ix - 1, N), xp1 = wrapi(ix + 1, N), xm2 = wrapi(ix - 2, N), xp2 = wrapi(ix + 2, N);
    const int ym1 = wrapi(iy - 1, N), yp1 = wrapi(iy + 1, N), ym2 = wrapi(iy - 2, N), yp2 = wrapi(iy + 2, N);
    const int zm1 = wrapi(iz - 1, N), zp1 = wrapi(iz + 1, N), zm2 = wrapi(iz - 2, N), zp2 = wrapi(iz + 2, N);
    const size_t rowyz = (size_t) iy * N + iz;
    const size_t rowxz = (size_t) ix * NN + iz;
    const size_t rowxy = (size_t) ix * NN + (size_t) iy * N;
    const double gx = 8.0 * (__ldg(pot + xp1 * NN + rowyz) - __ldg(pot + xm1 * NN + rowyz))
                    - (__ldg(pot + xp2 * NN + rowyz) - __ldg(pot + xm2 * NN + rowyz));
    const double gy = 8.0 * (__ldg(pot + rowxz + (size_t) yp1 * N) - __ldg(pot + rowxz + (size_t) ym1 * N))
                    - (__ldg(pot + rowxz + (size_t) yp2 * N) - __ldg(pot + rowxz + (size_t) ym2 * N));
    const double gz = 8.0 * (__ldg(pot + rowxy + zp1) - __ldg(pot + rowxy + zm1))
                    - (__ldg(pot + rowxy + zp2) - __ldg(pot + rowxy + zm2));
    fx[base] = -gx * inv12h;
    fy[base] = -gy * inv12h;
    fz[base] = -gz * inv12h;
}

// readout_potential / readout_force_{x,y,z} (gravpm.c:499-510) in one pass with the 4-point
// difference of the potential taken on the fly at each of the 8 CIC corners (the arithmetic of
// k_pm_gradient followed by k_pm_readout, value for value): no force meshes are written or
// read.  104 potential loads per particle; neighbouring particles share them through L1/L2.
__global__ void __launch_bounds__(128)
k_pm_readout_fused(const double *__restrict__ pos, const uint8_t *__restrict__ flags, int64_t n,
                   double cellsize, int N, const double *__restrict__ pot, double inv12h,
                   double *__restrict__ gravpm, double *__restrict__ potout)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double a0 = 0, a1 = 0, a2 = 0, p = 0;
    if(!(flags[i] & 3)) {
        int ic[3]; double res[3];
        cic_cell(pos + 3 * i, cellsize, ic, res);
        const double wx[2] = {1 - res[0], res[0]};
        const double wy[2] = {1 - res[1], res[1]};
        const double wz[2] = {1 - res[2], res[2]};
        const size_t NN = (size_t) N * N;
        // the 6 planes / rows / columns the two corners per axis need
        int xs[6], ys[6], zs[6];
#pragma unroll
        for(int k = 0; k < 6; k++) { xs[k] = wrapi(ic[0] - 2 + k, N); ys[k] = wrapi(ic[1] - 2 + k, N); zs[k] = wrapi(ic[2] - 2 + k, N); }
#pragma unroll
        for(int c = 0; c < 8; c++) {
            const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
            const int jx = 2 + ox, jy = 2 + oy, jz = 2 + oz;
            const size_t rowyz = (size_t) ys[jy] * N + zs[jz];
            const size_t rowxz = (size_t) xs[jx] * NN + zs[jz];
            const size_t rowxy = (size_t) xs[jx] * NN + (size_t) ys[jy] * N;
            const double gx = 8.0 * (__ldg(pot + xs[jx + 1] * NN + rowyz) - __ldg(pot + xs[jx - 1] * NN + rowyz))
                            - (__ldg(pot + xs[jx + 2] * NN + rowyz) - __ldg(pot + xs[jx - 2] * NN + rowyz));
            const double gy = 8.0 * (__ldg(pot + rowxz + (size_t) ys[jy + 1] * N) - __ldg(pot + rowxz + (size_t) ys[jy - 1] * N))
                            - (__ldg(pot + rowxz + (size_t) ys[jy + 2] * N) - __ldg(pot + rowxz + (size_t) ys[jy - 2] * N));
            const double gz = 8.0 * (__ldg(pot + rowxy + zs[jz + 1]) - __ldg(pot + rowxy + zs[jz - 1]))
                            - (__ldg(pot + rowxy + zs[jz + 2]) - __ldg(pot + rowxy + zs[jz - 2]));
            const double w = __dmul_rn(__dmul_rn(wx[ox], wy[oy]), wz[oz]);
            a0 = __dadd_rn(a0, __dmul_rn(w, -gx * inv12h));
            a1 = __dadd_rn(a1, __dmul_rn(w, -gy * inv12h));
            a2 = __dadd_rn(a2, __dmul_rn(w, -gz * inv12h));
            p  = __dadd_rn(p,  __dmul_rn(w, __ldg(pot + xs[jx] * NN + rowyz)));
        }
    }
    if(gravpm) { gravpm[3 * i] = a0; gravpm[3 * i + 1] = a1; gravpm[3 * i + 2] = a2; }
    if(potout) potout[i] = p;
}

// readout_potential / readout_force_{x,y,z} (gravpm.c:499-510) in one pass.
__global__ void __launch_bounds__(256)
k_pm_readout(const double *__restrict__ pos, const uint8_t *__restrict__ flags, int64_t n,
             double cellsize, int N,
             const double *__restrict__ pot, const double *__restrict__ fx,
             const double *__restrict__ fy, const double *__restrict__ fz,
             double *__restrict__ gravpm, double *__restrict__ potout)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double a0 = 0, a1 = 0, a2 = 0, p = 0;
    if(!(flags[i] & 3)) {
        int ic[3]; double res[3];
        cic_cell(pos + 3 * i, cellsize, ic, res);
        const int xs[2] = {wrapi(ic[0], N), wrapi(ic[0] + 1, N)};
        const int ys[2] = {wrapi(ic[1], N), wrapi(ic[1] + 1, N)};
        const int zs[2] = {wrapi(ic[2], N), wrapi(ic[2] + 1, N)};
        const double wx[2] = {1 - res[0], res[0]};
        const double wy[2] = {1 - res[1], res[1]};
        const double wz[2] = {1 - res[2], res[2]};
#pragma unroll
        for(int c = 0; c < 8; c++) {
            const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
            const double w = __dmul_rn(__dmul_rn(wx[ox], wy[oy]), wz[oz]);
            const size_t lin = ((size_t) xs[ox] * N + ys[oy]) * N + zs[oz];
            a0 = __dadd_rn(a0, __dmul_rn(w, __ldg(fx + lin)));
            a1 = __dadd_rn(a1, __dmul_rn(w, __ldg(fy + lin)));
            a2 = __dadd_rn(a2, __dmul_rn(w, __ldg(fz + lin)));
            p  = __dadd_rn(p,  __dmul_rn(w, __ldg(pot + lin)));
        }
    }
    if(gravpm) { gravpm[3 * i] = a0; gravpm[3 * i + 1] = a1; gravpm[3 * i + 2] = a2; }
    if(potout) potout[i] = p;
}

static double sinc_unnormed(double x)      // gravpm.c:295-302
{
    if(x < 1e-5 && x > -1e-5) {
        const double x2 = x * x;
        return 1.0 - x2 / 6. + x2 * x2 / 120.;
    }
    return sin(x) / x;
}

void pm_destroy(Engine *E)
{
    if(E->plans) {
        cufftDestroy(E->plan_fwd);
        cufftDestroy(E->plan_inv);
        E->plans = false;
    }
    pmfft_destroy(E);
    E->pm_rhok.release(); E->pm_table.release();
    E->mesh.release(); E->cplx.release(); E->fmesh.release(); E->ktab.release(); E->fftwork.release();
    E->Nmesh = 0;
}

// gravpm_init_periodic (gravpm.c:51-54) / petapm_init (petapm.c:104-223)
int pm_init(Engine *E, double Box, double Asmth, int Nmesh, double G)
{
    if(Nmesh < 8 || (Nmesh & 1)) return failmsg(E, "b200_pm_init: Nmesh must be even and >= 8");
    if(!(Box > 0)) return failmsg(E, "b200_pm_init: BoxSize must be positive");
    if((E->plans || E->ownfft) && E->Nmesh == Nmesh) {
        E->Box = Box; E->Asmth = Asmth; E->G = G;
    } else {
        pm_destroy(E);
        E->Box = Box; E->Asmth = Asmth; E->G = G; E->Nmesh = Nmesh;
        const size_t N = Nmesh, Nz = Nmesh / 2 + 1;
        CK(E->mesh.ensure(N * N * N));
        E->pm_fused = !(getenv("B200_PM_FUSED") && atoi(getenv("B200_PM_FUSED")) == 0);
        if(!E->pm_fused) CK(E->fmesh.ensure(3 * N * N * N));
        // the transforms: five shared-memory passes with the Green's function inside (pm_fft.cu) for mesh sizes
        // 2^a 3^b 5^c that fit a tile; cuFFT + k_pm_potential_transfer otherwise or with B200_PM_FFT=cufft
        const char *sel = getenv("B200_PM_FFT");
        if(pmfft_supported(Nmesh) && !(sel && !strcmp(sel, "cufft"))) {
            if(int rc = pmfft_init(E, Nmesh)) return rc;
            CK(E->cplx.ensure(pmfft_cplx_doubles(E)));
        } else {
        CK(E->cplx.ensure(2 * N * N * Nz));
        size_t ws_f = 0, ws_i = 0;
        CKF(cufftCreate(&E->plan_fwd));
        CKF(cufftCreate(&E->plan_inv));
        E->plans = true;
        CKF(cufftSetAutoAllocation(E->plan_fwd, 0));
        CKF(cufftSetAutoAllocation(E->plan_inv, 0));
        CKF(cufftMakePlan3d(E->plan_fwd, Nmesh, Nmesh, Nmesh, CUFFT_D2Z, &ws_f));
        CKF(cufftMakePlan3d(E->plan_inv, Nmesh, Nmesh, Nmesh, CUFFT_Z2D, &ws_i));
        const size_t ws = ws_f > ws_i ? ws_f : ws_i;
        CK(E->fftwork.ensure(ws + 256));
        CKF(cufftSetWorkArea(E->plan_fwd, E->fftwork.p));
        CKF(cufftSetWorkArea(E->plan_inv, E->fftwork.p));
        CKF(cufftSetStream(E->plan_fwd, E->stream));
        CKF(cufftSetStream(E->plan_inv, E->stream));
        }
    }
    // CIC deconvolution table, gravpm.c:403-407: tmp = (k*pi)/Nmesh; 1/sinc(tmp)^2.
    std::vector<double> tab(Nmesh);
    for(int i = 0; i < Nmesh; i++) {
        const int k = i <= Nmesh / 2 ? i : i - Nmesh;
        double tmp = (k * M_PI) / Nmesh;
        tmp = sinc_unnormed(tmp);
        tab[i] = 1. / (tmp * tmp);
    }
    CK(E->ktab.ensure(Nmesh));
    CK(cudaMemcpyAsync(E->ktab.p, tab.data(), Nmesh * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    E->potential_valid = false;
    E->NmeshWalk = Nmesh;
    return 0;
}

int pm_deposit(Engine *E)
{
    const size_t N = E->Nmesh;
    CK(cudaMemsetAsync(E->mesh.p, 0, N * N * N * sizeof(double), E->stream));
    if(E->n > 0) {
        const int bs = 256;
        const int64_t nb = (E->n + bs - 1) / bs;
        k_pm_deposit<<<(unsigned) nb, bs, 0, E->stream>>>(E->pos.p, E->mass.p, E->flags.p, E->n,
                                                        E->Box / E->Nmesh, E->Nmesh, E->mesh.p);
        CKL(E);
    }
    E->potential_valid = false;
    return 0;
}

int pm_cell_index(Engine *E, int32_t *d_icell)
{
    if(E->Nmesh == 0) return failmsg(E, "b200_pm_cell_index: call b200_pm_init first");
    if(E->n == 0) return 0;
    const int bs = 256;
    k_pm_cell_index<<<(unsigned) ((E->n + bs - 1) / bs), bs, 0, E->stream>>>(E->pos.p, E->n, E->Box / E->Nmesh, d_icell);
    CKL(E);
    return 0;
}

// The three force meshes from the potential mesh (parity hook and the unfused path).
int pm_force_meshes(Engine *E)
{
    const int N = E->Nmesh;
    const size_t N3 = (size_t) N * N * N;
    if(!E->potential_valid) return failmsg(E, "pm_force_meshes: no potential");
    CK(E->fmesh.ensure(3 * N3));
    double *fx = E->fmesh.p, *fy = fx + N3, *fz = fy + N3;
    const double h = E->Box / N;
    dim3 grid((N + 255) / 256, N, N);
    k_pm_gradient<<<grid, 256, 0, E->stream>>>(E->mesh.p, N, 1.0 / (12.0 * h), fx, fy, fz);
    CKL(E);
    E->fmesh_valid = true;
    return 0;
}

int pm_force(Engine *E, double *d_gravpm, double *d_pot)
{
    if(E->Nmesh == 0) return failmsg(E, "b200_pm_force: call b200_pm_init first");
    const int N = E->Nmesh, Nz = N / 2 + 1;
    timer_start(E, T_PM_DEPOSIT);
    if(int rc = pm_deposit(E)) return rc;
    timer_stop(E, T_PM_DEPOSIT);

    const double asmth2 = pow((2 * M_PI) * E->Asmth / N, 2);       // gravpm.c:386
    const double pot_factor = -E->G / (M_PI * E->Box);              // gravpm.c:392
    if(E->ownfft) {
        double *ps = nullptr;
        if(E->pm_power) {
            CK(E->pm_ps.ensure(3 * (size_t) N + 1));
            CK(cudaMemsetAsync(E->pm_ps.p, 0, (3 * (size_t) N + 1) * sizeof(double), E->stream));
            ps = E->pm_ps.p;
        }
        if(int rc = pmfft_potential(E, asmth2, pot_factor, (N - 1) / log(sqrt(3.) * N / 2.0), ps)) return rc;
        if(ps) E->pm_ps_valid = true;
    } else {
    timer_start(E, T_PM_FFT_FWD);
    CKF(cufftExecD2Z(E->plan_fwd, E->mesh.p, (cufftDoubleComplex *) E->cplx.p));
    E->launches += 1;
    timer_stop(E, T_PM_FFT_FWD);

    timer_start(E, T_PM_TRANSFER);
    {
        if(E->pm_power) {
            CK(E->pm_ps.ensure(3 * (size_t) N + 1));
            CK(cudaMemsetAsync(E->pm_ps.p, 0, (3 * (size_t) N + 1) * sizeof(double), E->stream));
            const double binsperunit = (N - 1) / log(sqrt(3.) * N / 2.0);       // gravpm.c:341 with size = Nmesh (gravpm.c:207)
            const size_t sm = 3 * (size_t) N * sizeof(double);
            CK(cudaFuncSetAttribute(k_pm_potential_transfer<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
            k_pm_potential_transfer<true><<<148 * 4, 256, sm, E->stream>>>((double2 *) E->cplx.p, N, Nz, E->ktab.p, asmth2, pot_factor,
                                                                          binsperunit, E->pm_ps.p);
            E->pm_ps_valid = true;
        } else
            k_pm_potential_transfer<false><<<148 * 16, 256, 0, E->stream>>>((double2 *) E->cplx.p, N, Nz, E->ktab.p, asmth2, pot_factor, 0.0, nullptr);
        CKL(E);
    }
    timer_stop(E, T_PM_TRANSFER);

    timer_start(E, T_PM_FFT_INV);
    CKF(cufftExecZ2D(E->plan_inv, (cufftDoubleComplex *) E->cplx.p, E->mesh.p));
    E->launches += 1;
    timer_stop(E, T_PM_FFT_INV);
    }
    E->potential_valid = true;

    E->fmesh_valid = false;
    const double h = E->Box / N;
    if(E->pm_fused) {
        timer_start(E, T_PM_GRADIENT); timer_stop(E, T_PM_GRADIENT);
        timer_start(E, T_PM_READOUT);
        if(E->n > 0 && (d_gravpm || d_pot)) {
            const int bs = 128;
            k_pm_readout_fused<<<(unsigned) ((E->n + bs - 1) / bs), bs, 0, E->stream>>>(E->pos.p, E->flags.p, E->n, h, N, E->mesh.p,
                                                                                    1.0 / (12.0 * h), d_gravpm, d_pot);
            CKL(E);
        }
        timer_stop(E, T_PM_READOUT);
        return 0;
    }
    timer_start(E, T_PM_GRADIENT);
    if(int rc = pm_force_meshes(E)) return rc;
    timer_stop(E, T_PM_GRADIENT);
    const size_t N3 = (size_t) N * N * N;
    double *fx = E->fmesh.p, *fy = fx + N3, *fz = fy + N3;
    timer_start(E, T_PM_READOUT);
    if(E->n > 0 && (d_gravpm || d_pot)) {
        const int bs = 256;
        k_pm_readout<<<(unsigned) ((E->n + bs - 1) / bs), bs, 0, E->stream>>>(E->pos.p, E->flags.p, E->n, h, N,
                                                                          E->mesh.p, fx, fy, fz, d_gravpm, d_pot);
        CKL(E);
    }
    timer_stop(E, T_PM_READOUT);
    return 0;
}

} // namespace b200
