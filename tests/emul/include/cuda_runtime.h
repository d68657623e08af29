/* cuda_runtime.h -- CPU stand-in for the CUDA runtime and the kernel language, enough to execute
 * mp-gadget_b200/csrc/steploop.cu (kernels AND host drivers, source unchanged) on the host.
 *
 * TEST INFRASTRUCTURE ONLY (tests/emul): it exists so that CUDA code written without access to a GPU
 * can be checked against the reference's golden vectors before its first hardware run.  Nothing in
 * the product links or loads this; the product path has no CPU mode.
 *
 * Execution model: a launch runs its blocks one after the other; the threads of a block are an
 * OpenMP team (one OS thread each), so __syncthreads() is a team barrier, __shared__ variables are
 * function statics, warp shuffles go through a team-wide exchange buffer.  "Device" memory is host
 * memory filled with a NaN pattern on allocation, so reads of uninitialised device memory show up. */
#ifndef EMUL_CUDA_RUNTIME_H
#define EMUL_CUDA_RUNTIME_H
#include <omp.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
using std::isfinite;

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); if(*p) memset(*p, 0xFF, n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if(n) memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for(size_t r = 0; r < height; r++) memmove((char *) d + r * dpitch, (const char *) s + r * spitch, width);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if(n) memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError(void) { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return 0; }

/* ---- kernel language ---- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
struct emul_dim { unsigned x, y, z; };
extern emul_dim emul_blockIdx, emul_blockDim, emul_gridDim;
struct emul_tid_t { unsigned x; };
static inline emul_tid_t emul_tid(void) { emul_tid_t t = {(unsigned) omp_get_thread_num()}; return t; }
#define threadIdx (emul_tid())
#define blockIdx emul_blockIdx
#define blockDim emul_blockDim
#define gridDim emul_gridDim
static inline void __syncthreads(void)
{
#pragma omp barrier
}
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r = {x, y}; return r; }
#define __align__(n)
/* dynamic shared memory: one static arena per launch (blocks run one after the other) */
extern unsigned char emul_dyn_smem[256 * 1024];
#define B200_DYN_SMEM(name) unsigned char *name = emul_dyn_smem
static inline double atomicAdd(double *p, double v)
{
    double old;
#pragma omp critical(emul_atomic_double)
    { old = *p; *p = old + v; }
    return old;
}
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long) v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicCAS(int *p, int expected, int desired)
{
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;          /* the value found, as CUDA returns it */
}
template <class T> static inline T __ldcg(const T *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while(v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while(v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
extern double emul_xchg[1024];
static inline double __shfl_xor_sync(unsigned, double v, int o)      /* every thread of the block takes part */
{
    const int t = omp_get_thread_num();
    emul_xchg[t] = v;
#pragma omp barrier
    const double r = emul_xchg[t ^ o];
#pragma omp barrier
    return r;
}

template <class F> static inline void emul_launch(unsigned grid, unsigned block, F body)
{
    emul_gridDim.x = grid; emul_blockDim.x = block;
    for(unsigned b = 0; b < grid; b++) {
        emul_blockIdx.x = b;
#pragma omp parallel num_threads(block)
        {
            if((unsigned) omp_get_num_threads() != block) { fprintf(stderr, "emul: could not start %u threads\n", block); abort(); }
            body();
        }
    }
}
/* tests/emul/build.py rewrites  kernel<<<grid, block, smem, stream>>>(args);  into this */
#define EMUL_LAUNCH(kernel, grid, block, ...) emul_launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#endif
