// Latency of the 8x8 diagonal-block factorisation (chol8 of csrc/tc_mcmc.cu) on one warp, alone and with busy neighbours
// (development aid).  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/chol8_bench scripts/chol8_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ bool chol8(double (&A)[8][8], double (&dinv)[8])
{
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = A[j][j];
        if (!(d > 0.0)) bad = true;
        double ri = (double)rsqrtf((float)d);
        const double hd = 0.5 * d;
        ri = ri * fma(-hd * ri, ri, 1.5);
        ri = ri * fma(-hd * ri, ri, 1.5);
        dinv[j] = ri;
        A[j][j] = d * ri;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) A[j][c] *= ri;
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
            for (int c = r; c < 8; ++c) A[r][c] = fma(-A[j][r], A[j][c], A[r][c]);
    }
    return bad;
}
// variant: the pivot's reciprocal square root by the double-precision intrinsic
__device__ __forceinline__ bool chol8_rsqrt(double (&A)[8][8], double (&dinv)[8])
{
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double d = A[j][j];
        if (!(d > 0.0)) bad = true;
        const double ri = rsqrt(d);
        dinv[j] = ri;
        A[j][j] = d * ri;
#pragma unroll
        for (int c = j + 1; c < 8; ++c) A[j][c] *= ri;
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
            for (int c = r; c < 8; ++c) A[r][c] = fma(-A[j][r], A[j][c], A[r][c]);
    }
    return bad;
}
__global__ void k(double *out, long long *cyc, int reps, int busy, int variant)
{
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        double A[8][8], dinv[8];
        double acc = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = i; j < 8; ++j) A[i][j] = (i == j ? 10.0 + i + 1e-3 * r : 0.1 * (i + j)) + acc * 1e-30;
            const bool bad = variant ? chol8_rsqrt(A, dinv) : chol8(A, dinv);
            acc += A[7][7] + dinv[3] + (bad ? 1 : 0);
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) { out[0] = acc; cyc[0] = (t1 - t0) / reps; }
    } else if (busy) {
        // neighbours hammering the FP64 pipe with independent FMAs (the trailing update of the factorisation)
        double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
        for (int i = 0; i < reps * 200; ++i) {
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
            a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
        }
        out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
}
int main()
{
    double *out; long long *cyc; cudaMallocManaged(&out, 8 * 1024); cudaMallocManaged(&cyc, 8);
    for (int variant = 0; variant < 2; ++variant)
        for (int busy = 0; busy < 2; ++busy)
            for (int warps : {1, 8, 16}) {
                if (!busy && warps > 1) continue;
                k<<<1, 32 * warps>>>(out, cyc, 200, busy, variant); cudaDeviceSynchronize();
                printf("chol8 %s, %2d warps (%s): %lld cycles per 8x8 block\n", variant ? "rsqrt()" : "rsqrtf + 2 Newton", warps, busy ? "neighbours busy on the FP64 pipe" : "alone", cyc[0]);
            }
    return 0;
}
