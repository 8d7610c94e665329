// Single-warp latency of ss_eval (development aid): cycles per evaluation, one warp resident.
#define TC_SS_PROFILE
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../transcriptioncycleinference_b200/csrc/tc_device.cuh"
using namespace tc;
__global__ void k(tc_construct C, int N, int algo, int reps, double *out, long long *cyc, int sumvec)
{
    SmemCell cv; Work w;
    int o = carve_cell(0, N, cv); o = carve_work(o, N, w);
    double *th = tc_smem + o; const int o_th = o; const int o_inc = o + 7 + N + 2;   /* interleaved (stage 1, stage 2) increments, all zero */
    const double d = 0.2521;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        tc_smem[cv.o_tg + i] = d * i; tc_smem[cv.o_dtg + i] = d; tc_smem[cv.o_ms2 + i] = (i % 3 == 0) ? NAN : 0.02 * i; tc_smem[cv.o_pp7 + i] = 0.05 * i;
        tc_smem[cv.o_iw + i] = 0.3; reinterpret_cast<int *>(tc_smem + cv.o_ik)[i] = min(i, N - 2);
    }
    cv.d = d;
    for (int i = threadIdx.x; i < 7 + N; i += blockDim.x) { th[i] = 0.0; tc_smem[o_inc + 2 * i] = 0.0; tc_smem[o_inc + 2 * i + 1] = 0.0; }
    __syncthreads();
    if (threadIdx.x == 0) { th[0] = 1.8; th[1] = 2.0; th[2] = 3.1; th[3] = 0.5; th[4] = 0.7; th[5] = 0.3; th[6] = 15.0; }
    for (int i = threadIdx.x; i < N; i += blockDim.x) th[7 + i] = sin(0.37 * i) * 3.0;
    __syncthreads();
    double acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) { th[1] = 2.0 + 1e-3 * (r & 7); acc += sumvec ? ss_eval(C, cv, SumVec{o_th, o_inc}, w, algo, false, nullptr, nullptr) : ss_eval(C, cv, SmemVec{o_th}, w, algo, false, nullptr, nullptr); __syncwarp(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = acc; cyc[0] = (t1 - t0) / reps; }
}
int main()
{
    tc_construct C{}; C.nsets = 1; C.L_ms2 = C.L_pp7 = 6.626; C.ms2_start[0] = 0.024; C.ms2_end[0] = 1.299; C.ms2_loopn[0] = 24;
    C.pp7_start[0] = 4.292; C.pp7_end[0] = 5.758; C.pp7_loopn[0] = 24;
    double *out; long long *cyc; cudaMallocManaged(&out, 8); cudaMallocManaged(&cyc, 8);
    for (int N : {120, 400}) for (int algo : {1}) for (int nt : {32}) for (int sumvec : {0, 1}) {
        size_t sm = sizeof(double) * (cell_doubles(N) + work_doubles(N) + 3 * (7 + N + 2) + 8);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        k<<<1, nt, sm>>>(C, N, algo, 200, out, cyc, sumvec); cudaDeviceSynchronize();
        long long pr[8]; cudaMemcpyFromSymbol(pr, tc_ss_prof, sizeof(pr)); long long z[8] = {0}; cudaMemcpyToSymbol(tc_ss_prof, z, sizeof(z));
        printf("sumvec=%d N=%d algo=%d threads=%d: %lld cycles/eval (ss=%g) phases: scan %lld tables %lld rows %lld resid %lld reduce %lld %s\n", sumvec, N, algo, nt, cyc[0], out[0] / 200,
               pr[0] / 200, pr[1] / 200, pr[2] / 200, pr[3] / 200, pr[4] / 200, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
