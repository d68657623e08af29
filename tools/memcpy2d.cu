// Microbenchmark: strided (2-D) DMA copies and zero-copy kernel access between a pinned
// 160-byte AoS on the host and device SoA, to choose the e2e ingest/write-back path.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/memcpy2d.cu -o /tmp/m2d && /tmp/m2d
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while(0)

__global__ void k_zc_read(const uint8_t *__restrict__ aos, int64_t n, double4 *__restrict__ a, double4 *__restrict__ b, double4 *__restrict__ c)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const double4 *r = (const double4 *) (aos + i * 160);
    a[i] = r[0];
    if(b) { b[i] = r[2]; c[i] = r[3]; }
}
__global__ void k_zc_write(uint8_t *__restrict__ aos, int64_t n, const double4 *__restrict__ a, const double2 *__restrict__ b)
{
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    double4 *r = (double4 *) (aos + i * 160);
    r[2] = a[i];
    double2 *q = (double2 *) (aos + i * 160 + 96);
    q[0] = b[i];
    *(double *) (aos + i * 160 + 152) = b[i].x;
}

int main()
{
    const int64_t n = 1 << 24;
    uint8_t *h, *d;
    CK(cudaMallocHost(&h, n * 160));
    memset(h, 1, n * 160);
    CK(cudaMalloc(&d, n * 160));
    uint8_t *soa;
    CK(cudaMalloc(&soa, n * 160));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for(int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0, s);
        CK(cudaMemcpyAsync(d, h, n * 160, cudaMemcpyHostToDevice, s));
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("bulk H2D 160 B/rec: %.2f ms  %.1f GB/s\n", ms, n * 160 / ms * 1e-6);
        cudaEventRecord(e0, s);
        CK(cudaMemcpyAsync(h, d, n * 160, cudaMemcpyDeviceToHost, s));
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("bulk D2H 160 B/rec: %.2f ms  %.1f GB/s\n", ms, n * 160 / ms * 1e-6);
        const int widths[4] = {32, 40, 48, 8};
        const int offs[4] = {0, 0, 64, 152};
        for(int w = 0; w < 4; w++) {
            cudaEventRecord(e0, s);
            CK(cudaMemcpy2DAsync(soa, widths[w], h + offs[w], 160, widths[w], n, cudaMemcpyHostToDevice, s));
            cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("2D H2D width %d: %.2f ms  payload %.1f GB/s\n", widths[w], ms, n * (double) widths[w] / ms * 1e-6);
            cudaEventRecord(e0, s);
            CK(cudaMemcpy2DAsync(h + offs[w], 160, soa, widths[w], widths[w], n, cudaMemcpyDeviceToHost, s));
            cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            printf("2D D2H width %d: %.2f ms  payload %.1f GB/s\n", widths[w], ms, n * (double) widths[w] / ms * 1e-6);
        }
        uint8_t *hd; CK(cudaHostGetDevicePointer((void **) &hd, h, 0));
        double4 *a = (double4 *) soa, *b = a + n, *c = b + n;
        cudaEventRecord(e0, s);
        k_zc_read<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(hd, n, a, nullptr, nullptr);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("zero-copy read 32 B/rec: %.2f ms  payload %.1f GB/s\n", ms, n * 32.0 / ms * 1e-6);
        cudaEventRecord(e0, s);
        k_zc_read<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(hd, n, a, b, c);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("zero-copy read 96 B/rec: %.2f ms  payload %.1f GB/s\n", ms, n * 96.0 / ms * 1e-6);
        cudaEventRecord(e0, s);
        k_zc_write<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(hd, n, a, (const double2 *) b);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("zero-copy write 56 B/rec: %.2f ms  payload %.1f GB/s\n", ms, n * 56.0 / ms * 1e-6);
        CK(cudaGetLastError());
    }
    // several strided copies at once on different streams (do the copy engines share the work?)
    {
        cudaStream_t st[4];
        for(int q = 0; q < 4; q++) CK(cudaStreamCreateWithFlags(&st[q], cudaStreamNonBlocking));
        for(int nsplit = 1; nsplit <= 4; nsplit *= 2) {
            for(int dir = 0; dir < 2; dir++) {
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0, st[0]);
                // two fields (40 B at 0, 48 B at 64), each split into nsplit row ranges -> 2*nsplit copies on up to 4 streams
                int k = 0;
                for(int f = 0; f < 2; f++) {
                    const int w = f ? 48 : 40, off = f ? 64 : 0;
                    for(int q = 0; q < nsplit; q++, k++) {
                        const int64_t a = n / nsplit * q, cnt = n / nsplit;
                        cudaStream_t s2 = st[k % 4];
                        if(dir == 0) CK(cudaMemcpy2DAsync(soa + (size_t) f * n * 48 + a * w, w, h + a * 160 + off, 160, w, cnt, cudaMemcpyHostToDevice, s2));
                        else CK(cudaMemcpy2DAsync(h + a * 160 + off, 160, soa + (size_t) f * n * 48 + a * w, w, w, cnt, cudaMemcpyDeviceToHost, s2));
                    }
                }
                for(int q = 1; q < 4; q++) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st[q]); cudaStreamWaitEvent(st[0], ev, 0); }
                cudaEventRecord(e1, st[0]); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                printf("%s 88 B/rec as %d strided copies on %d streams: %.2f ms  payload %.1f GB/s\n", dir ? "D2H" : "H2D", 2 * nsplit, 2 * nsplit < 4 ? 2 * nsplit : 4, ms, n * 88.0 / ms * 1e-6);
            }
        }
    }
    // bidirectional overlap: bulk H2D and D2H on two streams
    cudaStream_t s2; CK(cudaStreamCreate(&s2));
    uint8_t *h2; CK(cudaMallocHost(&h2, n * 160));
    cudaEventRecord(e0, s);
    CK(cudaMemcpyAsync(d, h, n * 160, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h2, soa, n * 160, cudaMemcpyDeviceToHost, s2));
    cudaStreamSynchronize(s2);
    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("bidirectional bulk: %.2f ms (each direction 2.68 GB)\n", ms);
    return 0;
}
