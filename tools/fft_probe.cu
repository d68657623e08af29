// Microbenchmark: 768^3 double-precision R2C as one 3-D cuFFT plan vs. three batched 1-D passes (z contiguous,
// then y and x strided), to decide how the PM step should factor its transforms.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fft_probe.cu -lcufft -o /tmp/fftp && /tmp/fftp [N]
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while(0)
#define CF(x) do { cufftResult r = (x); if(r != CUFFT_SUCCESS) { printf("%s: cufft error %d\n", #x, (int) r); return 1; } } while(0)
int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 768, NC = N / 2 + 1;
    double *re; cufftDoubleComplex *cx;
    CK(cudaMalloc(&re, sizeof(double) * (size_t) N * N * N));
    CK(cudaMalloc(&cx, sizeof(cufftDoubleComplex) * (size_t) N * N * NC));
    CK(cudaMemset(re, 0, sizeof(double) * (size_t) N * N * N));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    size_t ws;
    {
        cufftHandle p3; CF(cufftCreate(&p3)); CF(cufftMakePlan3d(p3, N, N, N, CUFFT_D2Z, &ws));
        printf("3-D plan work area %.2f GB\n", ws / 1e9);
        for(int r = 0; r < 3; r++) { cudaEventRecord(e0); CF(cufftExecD2Z(p3, re, cx)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("3-D D2Z: %.3f ms\n", ms); }
        cufftDestroy(p3);
    }
    {
        // z: N*N contiguous rows
        cufftHandle pz, py, px; int n[1] = {N};
        CF(cufftCreate(&pz)); CF(cufftMakePlanMany(pz, 1, n, NULL, 1, N, NULL, 1, NC, CUFFT_D2Z, N * N, &ws)); printf("z plan work %.2f GB\n", ws / 1e9);
        // y: for each x-plane, NC columns of stride NC, distance 1
        int emb[1] = {N};
        CF(cufftCreate(&py)); CF(cufftMakePlanMany(py, 1, n, emb, NC, 1, emb, NC, 1, CUFFT_Z2Z, NC, &ws)); printf("y plan work %.2f GB\n", ws / 1e9);
        // x: stride N*NC, N*NC columns
        CF(cufftCreate(&px)); CF(cufftMakePlanMany(px, 1, n, emb, N * NC, 1, emb, N * NC, 1, CUFFT_Z2Z, N * NC, &ws)); printf("x plan work %.2f GB\n", ws / 1e9);
        for(int r = 0; r < 3; r++) {
            float tz, ty, tx;
            cudaEventRecord(e0); CF(cufftExecD2Z(pz, re, cx)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tz, e0, e1);
            cudaEventRecord(e0);
            for(int i = 0; i < N; i++) CF(cufftExecZ2Z(py, cx + (size_t) i * N * NC, cx + (size_t) i * N * NC, CUFFT_FORWARD));
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ty, e0, e1);
            cudaEventRecord(e0); CF(cufftExecZ2Z(px, cx, cx, CUFFT_FORWARD)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tx, e0, e1);
            printf("1-D passes: z %.3f  y (N launches) %.3f  x %.3f  total %.3f ms\n", tz, ty, tx, tz + ty + tx);
        }
        // y+z as one batched 2-D plan, then x
        cufftHandle p2; int n2[2] = {N, N};
        CF(cufftCreate(&p2)); CF(cufftMakePlanMany(p2, 2, n2, NULL, 1, N * N, NULL, 1, N * NC, CUFFT_D2Z, N, &ws)); printf("2-D plan work %.2f GB\n", ws / 1e9);
        for(int r = 0; r < 3; r++) {
            float t2, tx;
            cudaEventRecord(e0); CF(cufftExecD2Z(p2, re, cx)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&t2, e0, e1);
            cudaEventRecord(e0); CF(cufftExecZ2Z(px, cx, cx, CUFFT_FORWARD)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&tx, e0, e1);
            printf("2-D batched + x: %.3f + %.3f = %.3f ms\n", t2, tx, t2 + tx);
        }
    }
    return 0;
}
