// Per-SM streaming bandwidth global -> shared (development aid): every CTA streams its own region of `bytes` bytes
// `reps` times (L2-resident after the first pass when 148 regions fit in L2) with (a) cp.async.bulk + mbarrier, DEPTH copies of
// CHUNK bytes in flight, (b) plain 16-byte loads, UNR per thread in flight.  Prints bytes/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ __align__(128) unsigned char smem[];
__device__ __forceinline__ unsigned su32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int DEPTH>
__global__ void __launch_bounds__(256) k_bulk(const char *g, size_t region, size_t stride_regions, int nreg, int chunk, int reps, long long *cyc, double *sink)
{
    __shared__ __align__(8) unsigned long long bar[DEPTH];
    const int tid = threadIdx.x;
    if (tid == 0) for (int i = 0; i < DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar + i)));
    __syncthreads();
    const int nchunks = (int)(region / chunk);
    double acc = 0;
    long long t0 = clock64();
    int gi = 0, gc = 0;
    for (int r = 0; r < reps; ++r) {
        const char *src = g + (size_t)((blockIdx.x + (size_t)r * stride_regions) % nreg) * region;
        // prologue
        int issued = 0;
        if (tid == 0) for (; issued < DEPTH - 1 && issued < nchunks; ++issued, ++gi) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar + gi % DEPTH)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(smem + (size_t)(gi % DEPTH) * chunk)), "l"(src + (size_t)issued * chunk), "r"(chunk), "r"(su32(bar + gi % DEPTH)) : "memory");
        }
        for (int c = 0; c < nchunks; ++c, ++gc) {
            __syncthreads();   // everyone done with the stage about to be overwritten
            if (tid == 0 && issued < nchunks) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar + gi % DEPTH)), "r"(chunk) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(smem + (size_t)(gi % DEPTH) * chunk)), "l"(src + (size_t)issued * chunk), "r"(chunk), "r"(su32(bar + gi % DEPTH)) : "memory");
                ++issued; ++gi;
            }
            unsigned ok; const unsigned par = (gc / DEPTH) & 1;
            do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(su32(bar + gc % DEPTH)), "r"(par) : "memory"); } while (!ok);
            acc += reinterpret_cast<const double *>(smem + (size_t)(gc % DEPTH) * chunk)[tid];
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 256 + tid] = acc;
}
template <int UNR>
__global__ void __launch_bounds__(256) k_ldg(const char *g, size_t region, size_t stride_regions, int nreg, int reps, long long *cyc, double *sink)
{
    const int tid = threadIdx.x;
    double acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const double2 *src = reinterpret_cast<const double2 *>(g + (size_t)((blockIdx.x + (size_t)r * stride_regions) % nreg) * region);
        const int n = (int)(region / 16);
        for (int e = tid; e + (UNR - 1) * 256 < n; e += UNR * 256) {
            double2 v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) v[u] = __ldcg(src + e + u * 256);
#pragma unroll
            for (int u = 0; u < UNR; ++u) acc += v[u].x + v[u].y;
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 256 + tid] = acc;
}
int main()
{
    const size_t region = 672 * 1024; const int nreg = 1184;
    char *g; cudaMalloc(&g, region * nreg); cudaMemset(g, 0, region * nreg);
    long long *cyc; double *sink; cudaMallocManaged(&cyc, 8 * 148); cudaMalloc(&sink, 8 * 148 * 256);
    const int reps = 16;
    auto report = [&](const char *name) { cudaDeviceSynchronize(); double m = 0; for (int i = 0; i < 148; ++i) m += cyc[i]; m /= 148; printf("%-46s %7.1f B/clk/SM (%.0f cycles per 672 KB)\n", name, region * reps / m, m / reps); };
    for (int pass = 0; pass < 2; ++pass) {
        const size_t stride = pass == 0 ? 0 : 148;   // 0: same region every rep (L2 hits), 148: a new region every rep (HBM)
        printf("--- %s\n", pass == 0 ? "L2-resident (each CTA re-reads its own 672 KB)" : "HBM (a new 672 KB region per pass)");
        cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        char nm[128];
        for (int chunk : {2048, 8192, 16384}) {
            k_bulk<4><<<148, 256, 4 * chunk>>>(g, region, stride, nreg, chunk, reps, cyc, sink); sprintf(nm, "bulk  chunk %5d x 3 in flight", chunk); report(nm);
            k_bulk<8><<<148, 256, 8 * chunk>>>(g, region, stride, nreg, chunk, reps, cyc, sink); sprintf(nm, "bulk  chunk %5d x 7 in flight", chunk); report(nm);
        }
        k_ldg<1><<<148, 256>>>(g, region, stride, nreg, reps, cyc, sink); report("ldg.128 x 1 per thread (4 KB in flight)");
        k_ldg<4><<<148, 256>>>(g, region, stride, nreg, reps, cyc, sink); report("ldg.128 x 4 per thread (16 KB in flight)");
        k_ldg<8><<<148, 256>>>(g, region, stride, nreg, reps, cyc, sink); report("ldg.128 x 8 per thread (32 KB in flight)");
        k_ldg<16><<<148, 256>>>(g, region, stride, nreg, reps, cyc, sink); report("ldg.128 x 16 per thread (64 KB in flight)");
    }
    return 0;
}
