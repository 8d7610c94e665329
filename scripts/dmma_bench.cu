// FP64 tensor-core (mma.sync m8n8k4) latency / throughput on one SM (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int CH> __global__ void k(double *out, long long *cyc, int n)
{
    double acc[CH][2];
    for (int c = 0; c < CH; ++c) { acc[c][0] = threadIdx.x; acc[c][1] = 1.0; }
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) dmma(acc[c][0], acc[c][1], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; ++c) s += acc[c][0] + acc[c][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH> __global__ void kf(double *out, long long *cyc, int n)
{
    double acc[CH];
    for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x;
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = fma(acc[c], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int c = 0; c < CH; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main()
{
    double *out; long long *cyc; cudaMallocManaged(&out, 8 * 4096); cudaMallocManaged(&cyc, 8 * 64);
    const int n = 2000;
    for (int warps : {1, 4, 8, 16}) {
        k<1><<<1, 32 * warps>>>(out, cyc, n); cudaDeviceSynchronize(); double c1 = (double)cyc[0] / n;
        k<4><<<1, 32 * warps>>>(out, cyc, n); cudaDeviceSynchronize(); double c4 = (double)cyc[0] / (4 * n);
        k<8><<<1, 32 * warps>>>(out, cyc, n); cudaDeviceSynchronize(); double c8 = (double)cyc[0] / (8 * n);
        kf<1><<<1, 32 * warps>>>(out, cyc, n); cudaDeviceSynchronize(); double f1 = (double)cyc[0] / n;
        kf<8><<<1, 32 * warps>>>(out, cyc, n); cudaDeviceSynchronize(); double f8 = (double)cyc[0] / (8 * n);
        printf("%2d warps on one SM: DMMA dependent %.1f cyc; 4 indep chains %.1f cyc/DMMA; 8 chains %.1f cyc/DMMA  (= %.1f FMA/clk/SM) | DFMA dep %.1f, 8 chains %.2f cyc/DFMA (= %.1f FMA/clk/SM)\n",
               warps, c1, c4, c8, warps * 256.0 / c8, f1, f8, warps * 32.0 / f8);
    }
    return 0;
}
