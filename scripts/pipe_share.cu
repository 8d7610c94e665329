// Development aid: does a stream of FP64 tensor-core MMAs (DMMA) on another warp of the same scheduler slow a dependent DFMA
// chain down as much as a stream of DFMAs does?  (The Cholesky factorisations have exactly this shape: one warp runs the serial
// chain of a diagonal block while the others update the trailing matrix.)  One CTA of 8 warps: warp 0 = the chain (timed),
// warps listed in `mask` = the stream, the rest idle.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *cyc, int n, int mode, unsigned mask)
{
    const int warp = threadIdx.x >> 5;
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    double s = 0.0;
    if (warp == 0) {
        double x = 1.0 + 1e-9 * threadIdx.x;
        const double a = 1.0 - 1e-12, b = 1e-12;
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) {
#pragma unroll
            for (int u = 0; u < 16; ++u) x = fma(x, a, b);
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) { cyc[0] = t1 - t0; done = 1; }
        s = x;
    } else if ((mask >> warp) & 1) {
        double acc[8][2];
        for (int c = 0; c < 8; ++c) { acc[c][0] = threadIdx.x; acc[c][1] = 1.0; }
        const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
        while (!done) {
            if (mode == 1) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) { acc[c][0] = fma(acc[c][0], a, b); acc[c][1] = fma(acc[c][1], a, b); }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) dmma(acc[c][0], acc[c][1], a, b);
            }
        }
        for (int c = 0; c < 8; ++c) s += acc[c][0] + acc[c][1];
    }
    out[threadIdx.x] = s;
}
int main()
{
    double *out; long long *cyc; cudaMallocManaged(&out, 8 * 256); cudaMallocManaged(&cyc, 8);
    const int n = 4000;
    k<<<1, 256>>>(out, cyc, n, 1, 0u); cudaDeviceSynchronize();
    printf("dependent DFMA chain alone: %.2f cycles per DFMA\n", (double)cyc[0] / (16.0 * n));
    const unsigned masks[4] = {1u << 4, 1u << 1, 0xfeu, 0xeeu};
    const char *names[4] = {"warp 4 (same scheduler)", "warp 1 (another scheduler)", "warps 1-7", "warps 1-3, 5-7 (not its scheduler partner)"};
    for (int m = 0; m < 4; ++m)
        for (int mode = 1; mode <= 2; ++mode) {
            k<<<1, 256>>>(out, cyc, n, mode, masks[m]); cudaDeviceSynchronize();
            printf("  with a %s stream on %-44s: %.2f cycles per DFMA of the chain\n", mode == 1 ? "DFMA" : "DMMA", names[m], (double)cyc[0] / (16.0 * n));
        }
    return 0;
}
