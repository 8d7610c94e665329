// FP64 latency micro-benchmarks on one warp (development aid): cycles per dependent op.
#include <cstdio>
#include <cuda_runtime.h>
#include "../transcriptioncycleinference_b200/csrc/tc_device.cuh"
using namespace tc;
#define RUN(name, ...)                                                       \
    {                                                                        \
        double x = seed + threadIdx.x * 1e-3; long long t0 = clock64();     \
        for (int i = 0; i < n; ++i) { __VA_ARGS__; }                                \
        long long t1 = clock64();                                            \
        if (threadIdx.x == 0) { out[k] = (double)(t1 - t0) / n; sink[k] = x; } \
        ++k;                                                                 \
    }
__global__ void lat(double *out, double *sink, double seed, int n)
{
    int k = 0;
    RUN("dfma", x = fma(x, 1.0000001, 1e-9));
    RUN("dadd", x = x + 1e-9);
    RUN("dmul", x = x * 1.0000001);
    RUN("floor", x = floor(x * 1.5) + 0.3);
    RUN("ddiv", x = 1.0 / (x + 1.5));
    RUN("dsqrt", x = sqrt(x + 2.0));
    RUN("log", x = log(x + 2.0));
    RUN("exp", x = exp(-x * 0.5));
    RUN("sincospi", { double s, c; sincospi(x, &s, &c); x = s + c; });
    RUN("philox", { u32x4 r = philox4x32_10((uint32_t)(x * 1e6), i, 3, 4, 5, 6); x = r.x * 1e-10; });
    RUN("normal_pair", { u32x4 r = philox4x32_10((uint32_t)(x * 1e6), i, 3, 4, 5, 6); double a, b; normal_pair(r, a, b); x = a + b; });
    RUN("chi2", x = chi2_draw(5, 7, i + (int)(x * 1e-9), 241.0) * 1e-3);
    RUN("ffma", { float f = (float)x; for (int j = 0; j < 8; ++j) f = fmaf(f, 1.0001f, 1e-6f); x = f; });
    RUN("shfl", x = __shfl_xor_sync(0xffffffffu, x, 1) + 1e-9);
    RUN("logf", { float f = __logf((float)x + 2.0f); x = f; });
}
int main()
{
    const char *names[] = {"dfma", "dadd", "dmul", "floor(+mul+add)", "ddiv(+add)", "dsqrt(+add)", "log(+add)", "exp(+mul)",
                           "sincospi(+add)", "philox(+cvt)", "philox+normal_pair", "chi2_draw", "8xffma+cvt", "shfl64+add", "__logf+cvt"};
    double *out, *sink;
    cudaMallocManaged(&out, 64 * 8); cudaMallocManaged(&sink, 64 * 8);
    for (int warps = 1; warps <= 4; warps *= 4) {
        lat<<<1, 32 * warps>>>(out, sink, 0.37, 2000);
        cudaDeviceSynchronize();
        printf("-- %d warp(s) in the CTA\n", warps);
        for (int i = 0; i < 15; ++i) printf("%-22s %8.1f cycles/iter\n", names[i], out[i]);
    }
    return 0;
}
