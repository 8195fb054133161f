// Micro-benchmark: DMMA.8x8x4 (mma.sync.m8n8k4.f64) issue rate and dependent-chain latency on this GPU, next to DADD.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_rate.bin scripts/micro/dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dmma(double *out, int iters, int chains)
{
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0 - threadIdx.x * 1e-3;
    double c[8][2] = {};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < chains) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dadd(double *out, int iters)
{
    double a = threadIdx.x * 1e-3 + 1.0;
    double c[8] = {};
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) c[k] += a;
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *d;
    cudaMalloc(&d, 148 * 1024 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev_clock = 0;
    cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int chains : {1, 2, 4, 8})
        for (int threads : {32, 128, 512, 1024})
        {
            k_dmma<<<148, threads>>>(d, 100, chains);
            cudaEventRecord(e0);
            k_dmma<<<148, threads>>>(d, iters, chains);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double per_sm = (double)iters * chains * (threads / 32);
            printf("DMMA chains/warp %d warps/SM %2d: %.3f ms, %.2f clk per DMMA per SM (at %d MHz), %.1f TFLOP/s\n", chains, threads / 32, ms,
                   ms * 1e-3 * dev_clock * 1e3 / per_sm, dev_clock / 1000, per_sm * 148 * 512 / (ms * 1e-3) / 1e12);
        }
    for (int threads : {128, 1024})
    {
        k_dadd<<<148, threads>>>(d, 100);
        cudaEventRecord(e0);
        k_dadd<<<148, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DADD warps/SM %2d: %.3f ms, %.2f thread-DADD per clk per SM\n", threads / 32, ms, (double)iters * 8 * threads / (ms * 1e-3 * dev_clock * 1e3));
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
