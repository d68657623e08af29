// pm_slab.cu -- the PM step on an x-slab of the mesh, for one rank of a
// multi-GPU run.  Replaces the reference's 2-D pencil decomposition + PFFT
// (petapm.c:127-150,584-885, pfft_execute_dft_* :305,344) by a 1-D slab
// decomposition: rank r owns mesh planes [x0, x0+nx) and the particles of its
// domain; the 3-D transform is   2-D cuFFT over (y,z) on the owned planes ->
// all-to-all transpose (done by the host harness over NCCL) -> 1-D cuFFT along x
// on a y-slab.  The Green's function is applied in the transposed layout, as
// PFFT_TRANSPOSED_OUT does in the reference (petapm.c:147-168).
//
// Real-space buffer:  planes [x0-halo, x0+nx+halo) x N x N  (density, then potential)
// Spectrum buffer:    [nx][N][N/2+1] complex (after the 2-D transforms)
// Transposed buffer:  [N (x)][nyloc][N/2+1] complex (y-slab, full x): exactly what the all-to-all delivers when
//                     rank s sends its block [ix][jy][kz] of my y-range -- the blocks, concatenated by source rank,
//                     ARE the x-major array, so the 1-D transforms along x run on the receive buffer in place
//                     (one strided cuFFT plan, stride nyloc*(N/2+1), batch nyloc*(N/2+1)) and nothing is unpacked.
// Forces are the 4-point difference of the potential fused into the CIC readout
// (see pm.cu header for why that equals the reference's k-space gradient).
#include "pm_slab.h"
#include <math.h>

namespace b200 {

static const char *cufft_str2(cufftResult r) { return r == CUFFT_ALLOC_FAILED ? "CUFFT_ALLOC_FAILED" : (r == CUFFT_INVALID_SIZE ? "CUFFT_INVALID_SIZE" : "CUFFT_ERROR"); }
#define CKF(call) do { cufftResult _r = (call); if(_r != CUFFT_SUCCESS) return failmsg(E, std::string(#call) + ": " + cufft_str2(_r)); } while(0)

__device__ __forceinline__ int wrapi2(int i, int N) { i %= N; return i < 0 ? i + N : i; }

__device__ __forceinline__ void cic_cell2(const double *__restrict__ p, double cellsize, int ic[3], double res[3])
{
#pragma unroll
    for(int k = 0; k < 3; k++) {
        const double t = __ddiv_rn(p[k], cellsize);      // petapm.c:976-980
        const double f = floor(t);
        ic[k] = (int) f;
        res[k] = t - f;
    }
}

// local plane index of global plane gx, or -1 when outside [x0-halo, x0+nx+halo)
__device__ __forceinline__ int local_plane(int gx, int x0, int nxh, int halo, int N)
{
    const int d = wrapi2(gx - x0 + halo, N);
    return d < nxh ? d : -1;
}

__global__ void __launch_bounds__(256)
k_slab_deposit(const double *__restrict__ pos, const float *__restrict__ mass, const uint8_t *__restrict__ flags,
               int64_t n, double cellsize, int N, int x0, int nx, int halo, double *__restrict__ mesh, int *__restrict__ err)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(flags[i] & 3) return;
    int ic[3]; double res[3];
    cic_cell2(pos + 3 * i, cellsize, ic, res);
    const double m = (double) mass[i];
    const int nxh = nx + 2 * halo;
    const int xs[2] = {local_plane(ic[0], x0, nxh, halo, N), local_plane(ic[0] + 1, x0, nxh, halo, N)};
    if(xs[0] < 0 || xs[1] < 0) { atomicAdd(err, 1); return; }
    const int ys[2] = {wrapi2(ic[1], N), wrapi2(ic[1] + 1, N)};
    const int zs[2] = {wrapi2(ic[2], N), wrapi2(ic[2] + 1, N)};
    const double wx[2] = {1 - res[0], res[0]}, wy[2] = {1 - res[1], res[1]}, wz[2] = {1 - res[2], res[2]};
#pragma unroll
    for(int c = 0; c < 8; c++) {
        const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
        const double w = __dmul_rn(__dmul_rn(__dmul_rn(wx[ox], wy[oy]), wz[oz]), m);
        atomicAdd(mesh + ((size_t) xs[ox] * N + ys[oy]) * N + zs[oz], w);
    }
}

// potential_transfer (gravpm.c:383-454) in the transposed layout [ix][jy][iz], ky = y0 + jy
__global__ void __launch_bounds__(256)
k_slab_transfer(double2 *__restrict__ v, int N, int Nz, int y0, int ny, const double *__restrict__ ktab,
                double asmth2, double pot_factor)
{
    const size_t total = (size_t) ny * N * Nz;
    for(size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t) gridDim.x * blockDim.x) {
        const int iz = (int) (idx % Nz);
        const size_t row = idx / Nz;
        const int iy = y0 + (int) (row % ny);
        const int ix = (int) (row / ny);
        const int kx = ix <= N / 2 ? ix : ix - N;
        const int ky = iy <= N / 2 ? iy : iy - N;
        const long long k2 = (long long) kx * kx + (long long) ky * ky + (long long) iz * iz;
        double2 val = v[idx];
        if(k2 == 0) { val.x = 0.0; val.y = 0.0; }
        else {
            const double smth = exp((double) (-k2) * asmth2) / (double) k2;
            const double f = (ktab[ix] * ktab[iy]) * ktab[iz];
            const double fac = ((pot_factor * smth) * f) * f;
            val.x *= fac; val.y *= fac;
        }
        v[idx] = val;
    }
}

// readout_potential / readout_force_* (gravpm.c:499-510) with the 4-point
// difference of the potential evaluated on the fly at each CIC corner (the arithmetic of
// k_pm_readout_fused, pm.cu, on the slab's local planes).  The six planes / rows / columns the
// two corners per axis need are wrapped once per particle.
__global__ void __launch_bounds__(128)
k_slab_readout(const double *__restrict__ pos, const uint8_t *__restrict__ flags, int64_t n, double cellsize,
               int N, int x0, int nx, int halo, const double *__restrict__ pot, double inv12h,
               double *__restrict__ gravpm, double *__restrict__ potout, int *__restrict__ err)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double a0 = 0, a1 = 0, a2 = 0, p = 0;
    if(!(flags[i] & 3)) {
        int ic[3]; double res[3];
        cic_cell2(pos + 3 * i, cellsize, ic, res);
        const int nxh = nx + 2 * halo;
        const double wx[2] = {1 - res[0], res[0]}, wy[2] = {1 - res[1], res[1]}, wz[2] = {1 - res[2], res[2]};
        const size_t NN = (size_t) N * N;
        int xs[6], ys[6], zs[6];
        // local plane of global plane ic[0] - 2; the following five are consecutive (no wrap inside slab + halos)
        const int lx0 = local_plane(ic[0] - 2, x0, nxh, halo, N);
        const int y0 = wrapi2(ic[1] - 2, N), z0 = wrapi2(ic[2] - 2, N);
        const bool bad = lx0 < 0 || lx0 + 5 >= nxh;
#pragma unroll
        for(int k = 0; k < 6; k++) {
            xs[k] = bad ? 0 : lx0 + k;
            ys[k] = y0 + k >= N ? y0 + k - N : y0 + k;
            zs[k] = z0 + k >= N ? z0 + k - N : z0 + k;
        }
        if(bad) atomicAdd(err, 1);
        else {
#pragma unroll
            for(int c = 0; c < 8; c++) {
                const int ox = c & 1, oy = (c >> 1) & 1, oz = (c >> 2) & 1;
                const int jx = 2 + ox, jy = 2 + oy, jz = 2 + oz;
                const size_t rowyz = (size_t) ys[jy] * N + zs[jz];
                const size_t rowxz = (size_t) xs[jx] * NN + zs[jz];
                const size_t rowxy = (size_t) xs[jx] * NN + (size_t) ys[jy] * N;
                const double gx_ = 8.0 * (__ldg(pot + xs[jx + 1] * NN + rowyz) - __ldg(pot + xs[jx - 1] * NN + rowyz))
                                 - (__ldg(pot + xs[jx + 2] * NN + rowyz) - __ldg(pot + xs[jx - 2] * NN + rowyz));
                const double gy_ = 8.0 * (__ldg(pot + rowxz + (size_t) ys[jy + 1] * N) - __ldg(pot + rowxz + (size_t) ys[jy - 1] * N))
                                 - (__ldg(pot + rowxz + (size_t) ys[jy + 2] * N) - __ldg(pot + rowxz + (size_t) ys[jy - 2] * N));
                const double gz_ = 8.0 * (__ldg(pot + rowxy + zs[jz + 1]) - __ldg(pot + rowxy + zs[jz - 1]))
                                 - (__ldg(pot + rowxy + zs[jz + 2]) - __ldg(pot + rowxy + zs[jz - 2]));
                const double w = __dmul_rn(__dmul_rn(wx[ox], wy[oy]), wz[oz]);
                a0 = __dadd_rn(a0, __dmul_rn(w, -gx_ * inv12h));
                a1 = __dadd_rn(a1, __dmul_rn(w, -gy_ * inv12h));
                a2 = __dadd_rn(a2, __dmul_rn(w, -gz_ * inv12h));
                p  = __dadd_rn(p,  __dmul_rn(w, __ldg(pot + xs[jx] * NN + rowyz)));
            }
        }
    }
    if(gravpm) { gravpm[3 * i] = a0; gravpm[3 * i + 1] = a1; gravpm[3 * i + 2] = a2; }
    if(potout) potout[i] = p;
}

static double sinc_unnormed2(double x)      // gravpm.c:295-302
{
    if(x < 1e-5 && x > -1e-5) { const double x2 = x * x; return 1.0 - x2 / 6. + x2 * x2 / 120.; }
    return sin(x) / x;
}

void pmslab_destroy(Engine *E)
{
    SlabPM *S = E->slab;
    if(!S) return;
    if(S->plans) { cufftDestroy(S->p2f); cufftDestroy(S->p2i); cufftDestroy(S->p1); }
    S->real.release(); S->cplx.release(); S->cplxT.release(); S->ktab.release(); S->err.release(); S->work.release();
    delete S;
    E->slab = nullptr;
}

int pmslab_init(Engine *E, double Box, double Asmth, int Nmesh, double G, int rank, int nranks, int halo,
                void **real_buf, void **cplx_buf, void **cplxT_buf)
{
    if(Nmesh % (2 * nranks) != 0) return failmsg(E, "b200_pmslab_init: Nmesh must be a multiple of 2*nranks");
    if(halo < 4) return failmsg(E, "b200_pmslab_init: halo must be >= 4 planes (CIC + 4-point difference + domain/mesh offset)");
    pmslab_destroy(E);
    SlabPM *S = new SlabPM();
    E->slab = S;
    S->Box = Box; S->Asmth = Asmth; S->G = G; S->N = Nmesh; S->Nz = Nmesh / 2 + 1;
    S->rank = rank; S->nranks = nranks; S->halo = halo;
    S->nx = Nmesh / nranks; S->x0 = rank * S->nx; S->ny = S->nx; S->y0 = S->x0;
    if(S->nx + 2 * halo > Nmesh && nranks > 1) return failmsg(E, "b200_pmslab_init: slab + halos exceed the mesh");
    const size_t N = Nmesh, Nz = S->Nz;
    CK(S->real.ensure((size_t) (S->nx + 2 * halo) * N * N));
    CK(S->cplx.ensure(2 * (size_t) S->nx * N * Nz));
    CK(S->cplxT.ensure(2 * (size_t) S->ny * N * Nz));
    CK(S->err.ensure(4));
    CK(cudaMemsetAsync(S->err.p, 0, 4 * sizeof(int), E->stream));
    // three plans that never run concurrently share one work area
    int n2[2] = {Nmesh, Nmesh};
    int n1[1] = {Nmesh};
    int emb[1] = {Nmesh};
    const int strideT = S->ny * (int) Nz;
    size_t ws[3] = {0, 0, 0};
    CKF(cufftCreate(&S->p2f)); CKF(cufftCreate(&S->p2i)); CKF(cufftCreate(&S->p1));
    S->plans = true;
    CKF(cufftSetAutoAllocation(S->p2f, 0)); CKF(cufftSetAutoAllocation(S->p2i, 0)); CKF(cufftSetAutoAllocation(S->p1, 0));
    CKF(cufftMakePlanMany(S->p2f, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, S->nx, &ws[0]));
    CKF(cufftMakePlanMany(S->p2i, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, S->nx, &ws[1]));
    CKF(cufftMakePlanMany(S->p1, 1, n1, emb, strideT, 1, emb, strideT, 1, CUFFT_Z2Z, strideT, &ws[2]));
    size_t wmax = ws[0] > ws[1] ? ws[0] : ws[1]; if(ws[2] > wmax) wmax = ws[2];
    CK(S->work.ensure(wmax + 256));
    CKF(cufftSetWorkArea(S->p2f, S->work.p)); CKF(cufftSetWorkArea(S->p2i, S->work.p)); CKF(cufftSetWorkArea(S->p1, S->work.p));
    CKF(cufftSetStream(S->p2f, E->stream)); CKF(cufftSetStream(S->p2i, E->stream)); CKF(cufftSetStream(S->p1, E->stream));
    std::vector<double> tab(Nmesh);
    for(int i = 0; i < Nmesh; i++) {
        const int k = i <= Nmesh / 2 ? i : i - Nmesh;
        double tmp = (k * M_PI) / Nmesh;
        tmp = sinc_unnormed2(tmp);
        tab[i] = 1. / (tmp * tmp);
    }
    CK(S->ktab.ensure(Nmesh));
    CK(cudaMemcpyAsync(S->ktab.p, tab.data(), Nmesh * sizeof(double), cudaMemcpyHostToDevice, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    // the short-range walk needs the same mesh scalars
    E->Box = Box; E->Asmth = Asmth; E->G = G; E->NmeshWalk = Nmesh;
    if(real_buf) *real_buf = S->real.p;
    if(cplx_buf) *cplx_buf = S->cplx.p;
    if(cplxT_buf) *cplxT_buf = S->cplxT.p;
    return 0;
}

int pmslab_deposit(Engine *E, int64_t n_own, bool check)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "b200_pmslab_deposit: call b200_pmslab_init first");
    if(n_own > E->n) return failmsg(E, "b200_pmslab_deposit: n_own exceeds the particle count");
    const size_t N = S->N;
    timer_start(E, T_PM_DEPOSIT);
    CK(cudaMemsetAsync(S->real.p, 0, (size_t) (S->nx + 2 * S->halo) * N * N * sizeof(double), E->stream));
    if(n_own > 0) {
        k_slab_deposit<<<(unsigned) ((n_own + 255) / 256), 256, 0, E->stream>>>(E->pos.p, E->mass.p, E->flags.p, n_own, S->Box / S->N,
                                                                            S->N, S->x0, S->nx, S->halo, S->real.p, S->err.p);
        CKL(E);
    }
    timer_stop(E, T_PM_DEPOSIT);
    if(!check) return 0;
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, S->err.p, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(herr) { cudaMemsetAsync(S->err.p, 0, sizeof(int), E->stream); return failmsg(E, "b200_pmslab_deposit: particles outside this rank's slab + halo"); }
    return 0;
}

int pmslab_fft2d(Engine *E, int inverse)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "b200_pmslab_fft2d: no slab");
    double *owned = S->real.p + (size_t) S->halo * S->N * S->N;
    timer_start(E, inverse ? T_PM_FFT_INV : T_PM_FFT_FWD);
    if(!inverse) CKF(cufftExecD2Z(S->p2f, owned, (cufftDoubleComplex *) S->cplx.p));
    else CKF(cufftExecZ2D(S->p2i, (cufftDoubleComplex *) S->cplx.p, owned));
    E->launches += 2;
    timer_stop(E, inverse ? T_PM_FFT_INV : T_PM_FFT_FWD);
    return 0;
}

int pmslab_fft1d(Engine *E, int inverse)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "b200_pmslab_fft1d: no slab");
    cufftDoubleComplex *p = (cufftDoubleComplex *) S->cplxT.p;
    CKF(cufftExecZ2Z(S->p1, p, p, inverse ? CUFFT_INVERSE : CUFFT_FORWARD));
    E->launches += 1;
    return 0;
}

int pmslab_transfer(Engine *E)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "b200_pmslab_transfer: no slab");
    timer_start(E, T_PM_TRANSFER);
    const double asmth2 = pow((2 * M_PI) * S->Asmth / S->N, 2);
    const double pot_factor = -S->G / (M_PI * S->Box);
    k_slab_transfer<<<148 * 8, 256, 0, E->stream>>>((double2 *) S->cplxT.p, S->N, S->Nz, S->y0, S->ny, S->ktab.p, asmth2, pot_factor);
    CKL(E);
    timer_stop(E, T_PM_TRANSFER);
    return 0;
}

int pmslab_readout(Engine *E, int64_t n_own, double *d_gravpm, double *d_pot, bool check)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "b200_pmslab_readout: no slab");
    timer_start(E, T_PM_READOUT);
    if(n_own > 0) {
        const double h = S->Box / S->N;
        k_slab_readout<<<(unsigned) ((n_own + 127) / 128), 128, 0, E->stream>>>(E->pos.p, E->flags.p, n_own, h, S->N, S->x0, S->nx, S->halo,
                                                                            S->real.p, 1.0 / (12.0 * h), d_gravpm, d_pot, S->err.p);
        CKL(E);
    }
    timer_stop(E, T_PM_READOUT);
    if(!check) return 0;
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, S->err.p, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CK(cudaStreamSynchronize(E->stream));
    if(herr) { cudaMemsetAsync(S->err.p, 0, sizeof(int), E->stream); return failmsg(E, "b200_pmslab_readout: particles outside this rank's slab + halo"); }
    return 0;
}

// Set the stream of the three plans (the PM step may run on a side stream).
int pmslab_set_stream(Engine *E, cudaStream_t st)
{
    SlabPM *S = E->slab;
    if(!S) return failmsg(E, "pmslab_set_stream: no slab");
    CKF(cufftSetStream(S->p2f, st)); CKF(cufftSetStream(S->p2i, st)); CKF(cufftSetStream(S->p1, st));
    return 0;
}

} // namespace b200
