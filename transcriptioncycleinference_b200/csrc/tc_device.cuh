// tc_device.cuh — device-side building blocks of libtcmcmc (sm_100a).
//
//  * Philox4x32-10 counter-based RNG, Box-Muller normals, Marsaglia-Tsang chi-square
//  * CTA-wide reductions on warp shuffles
//  * the forward model + residual sum of squares for one (cell, theta), evaluated cooperatively by
//    one CTA out of shared memory.  Behavioural spec: SURVEY.md Appendix B.2, i.e.
//      src/SumofSquaresFunction_TranscriptionCycleMCMC.m:28-64,
//      src/dependencies/ConstantElongationSim.m:33-67,
//      src/GetFluorFromPolPos.m:18-70            (paths relative to /root/reference).
//    Not a translation: the reference builds an m x n position matrix and makes ~10 masked passes
//    over it; here polymerases loaded in the same step form one cohort of integer size
//    n_i = floor(c_i) - floor(c_{i-1}) at position v*(t_j - t_i), so a row of the matrix collapses
//    to a (cohort, time) sum — evaluated either pair by pair (TC_ALGO_PAIRS) or, on the uniform
//    t_interp grid, through the banded structure of the per-lag response (TC_ALGO_TOEPLITZ).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/tcmcmc.h"

namespace tc {

// ------------------------------------------------------------------------------------------ RNG
struct u32x4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2,
                                                      uint32_t &c3, uint32_t k0, uint32_t k1)
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

// Stream kinds: which draw of MCMC step `step` a counter addresses.
enum { RK_Z1 = 0, RK_Z2 = 1, RK_U = 2, RK_CHI2 = 3 };

// counter = (slot, step, uid_lo, uid_hi[23:0] << 8 | kind); key = seed
__device__ __forceinline__ u32x4 draw(uint64_t seed, uint64_t uid, uint32_t step, uint32_t kind,
                                      uint32_t slot)
{
    return philox4x32_10(slot, step, (uint32_t)uid, ((uint32_t)(uid >> 32) << 8) | kind,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 53-bit uniform in (0,1) from two 32-bit words
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo)
{
    const uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

// two standard normals from one Philox block (Box-Muller, FP64)
__device__ __forceinline__ void normal_pair(const u32x4 &r, double &z0, double &z1)
{
    const double ua = u01(r.x, r.y), ub = u01(r.z, r.w);
    const double rad = sqrt(-2.0 * log(ua));
    double s, c;
    sincospi(2.0 * ub, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// chi-square(dof) = 2*Gamma(dof/2) by Marsaglia-Tsang (dof >= 2); attempt t uses slots 2t, 2t+1.
__device__ inline double chi2_draw(uint64_t seed, uint64_t uid, uint32_t step, double dof)
{
    const double a = 0.5 * dof;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0; t < 64; ++t) {
        const u32x4 r0 = draw(seed, uid, step, RK_CHI2, 2 * t);
        double x, unused;
        normal_pair(r0, x, unused);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const u32x4 r1 = draw(seed, uid, step, RK_CHI2, 2 * t + 1);
        const double u = u01(r1.x, r1.y);
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return 2.0 * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return 2.0 * d * v;
    }
    return 2.0 * d;   // unreachable in practice (acceptance > 0.95 per attempt)
}

// ------------------------------------------------------------------------------- reductions
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of (a, b) over the CTA, result in every thread.  red: >= 2*32 doubles of shared memory.
__device__ __forceinline__ void block_sum2(double &a, double &b, double *red)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();                      // red[] free to overwrite
    if ((threadIdx.x & 31) == 0) { red[2 * w] = a; red[2 * w + 1] = b; }
    __syncthreads();
    double sa = 0, sb = 0;
    for (int i = 0; i < nw; ++i) { sa += red[2 * i]; sb += red[2 * i + 1]; }
    a = sa; b = sb;
}

// ------------------------------------------------------------------------- shared-memory views
struct CellView {        // one cell's constants, staged in shared memory
    int N;
    double d;            // mean(diff(t))                       SumofSquares...m:29
    double *tg;          // grid the model runs on (t_interp, or raw t)  [N]
    double *dtg;         // tg[i+1]-tg[i]                        [N]   ConstantElongationSim.m:43-45
    double *ms2, *pp7;   // data, NaN = missing                  [N]
    double *iw;          // interp1 weight of experimental time j [N]
    int *ik;             // interp1 bracketing index (-1: outside) [N]
};
struct Work {            // per-evaluation scratch in shared memory, each [N+1]
    double *rd, *K, *G1, *G2, *F1, *F2;
    int *thr;            // 8 lag thresholds
    double *red;         // 64 doubles for reductions
};

__host__ __device__ inline int work_doubles(int N) { return 6 * (N + 2) + 64 + 4; }
__host__ __device__ inline int cell_doubles(int N) { return 5 * (N + 1) + (N + 2) / 2 + 1; }

__device__ inline void carve_cell(double *&p, int N, CellView &cv)
{
    cv.N = N;
    cv.tg = p; p += N + 1;
    cv.dtg = p; p += N + 1;
    cv.ms2 = p; p += N + 1;
    cv.pp7 = p; p += N + 1;
    cv.iw = p; p += N + 1;
    cv.ik = reinterpret_cast<int *>(p); p += (N + 2) / 2 + 1;
}
__device__ inline void carve_work(double *&p, int N, Work &w)
{
    w.rd = p; p += N + 2;
    w.K = p; p += N + 2;
    w.G1 = p; p += N + 2;
    w.G2 = p; p += N + 2;
    w.F1 = p; p += N + 2;
    w.F2 = p; p += N + 2;
    w.red = p; p += 64;
    w.thr = reinterpret_cast<int *>(p); p += 4;
}

struct CellsDev {        // device-resident packed dataset (one per device)
    int ncells;
    const int *N;
    const long long *off;
    const double *t, *tg, *dtraw, *dtg, *ms2, *pp7, *iw, *dmean;
    const int *ik;
};

// stage cell `cid` into shared memory; raw_grid selects the raw experimental times as model grid
__device__ inline void load_cell(const CellsDev &cd, int cid, bool raw_grid, CellView &cv)
{
    const int N = cv.N;
    const long long o = cd.off[cid];
    const double *g = raw_grid ? cd.t : cd.tg;
    const double *dg = raw_grid ? cd.dtraw : cd.dtg;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        cv.tg[i] = g[o + i];
        cv.dtg[i] = dg[o + i];
        cv.ms2[i] = cd.ms2[o + i];
        cv.pp7[i] = cd.pp7[o + i];
        cv.iw[i] = cd.iw[o + i];
        cv.ik[i] = cd.ik[o + i];
    }
    cv.d = cd.dmean[cid];
}

// ------------------------------------------------------------------------- forward model pieces

// literal response of one loop set at position p  (GetFluorFromPolPos.m:49-52 / :61-64)
__device__ __forceinline__ double loop_response(double p, double s, double e, double L, double fv)
{
    double val = 0.0;
    if (p > e && p < L) val = fv;
    if (p > s && p < e) val = (p - s) * fv / (e - s);
    return val;
}

// Loaded-polymerase counts.  K[0] = 0, K[i+1] = floor(counter after step i)
// (ConstantElongationSim.m:53-61).  The running sum is taken in the reference's order with
// separately rounded products (no FMA contraction), because floor() is discontinuous.
// Fast path: warp 0 scans in parallel; if any partial sum lands within 1e-7 of an integer — where a
// different association could flip a floor — lane 0 redoes the scan sequentially.  Both paths give
// the same integers as the sequential reference order.
__device__ inline void scan_counts(int N, const double *rd, double *K, bool force_sequential)
{
    const int lane = threadIdx.x & 31;               // called by ONE full warp
    const int n = N - 1;                       // increments rd[0..n-1]
    bool redo = force_sequential;
    if (!force_sequential) {
        const int chunk = (n + 31) >> 5;
        const int b = lane * chunk, e = min(b + chunk, n);
        double loc = 0.0;
        for (int i = b; i < e; ++i) loc = __dadd_rn(loc, rd[i]);
        double incl = loc;                     // inclusive scan of chunk sums over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = __dadd_rn(incl, up);
        }
        double c = __dadd_rn(incl, -loc);      // exclusive prefix for this lane
        bool risky = false;
        for (int i = b; i < e; ++i) {
            c = __dadd_rn(c, rd[i]);
            const double f = floor(c);
            // c == 0: every increment so far is exactly 0 (they are all >= 0) -> exact in any order
            risky |= (c != 0.0) && ((c - f < 1e-7) || (f + 1.0 - c < 1e-7));
            K[i + 1] = f;
        }
        redo = __any_sync(0xffffffffu, risky);
        if (lane == 0) K[0] = 0.0;
    }
    if (redo && lane == 0) {
        double c = 0.0;
        K[0] = 0.0;
        for (int i = 0; i < n; ++i) {
            c = __dadd_rn(c, rd[i]);
            K[i + 1] = floor(c);
        }
    }
}

// smallest lag in [1, N] with v*(d*lag) > x (strict) or >= x; N when none
__device__ inline int first_lag(double v, double d, double x, int N, bool strict)
{
    const double vd = v * d;
    if (!(vd > 0.0)) return N;
    double q = floor(x / vd);
    int g = q < 1.0 ? 1 : (q > (double)N ? N : (int)q);
    auto pred = [&](int lag) {
        const double p = v * (d * (double)lag);
        return strict ? (p > x) : (p >= x);
    };
    while (g > 1 && pred(g - 1)) --g;
    while (g < N && !pred(g)) ++g;
    return g;
}

// The residual sum of squares of one (cell, theta) — SumofSquares...m:1-65 — by the whole CTA.
// th: theta (shared or global).  On return every thread holds SS.  When out1/out2 != nullptr the
// model curves [A*MS2, PP7] on the model grid are also written there (tc_forward).
// All threads of the CTA must call this.
__device__ inline double ss_eval(const tc_construct &C, const CellView &cv, const double *th,
                                 Work &w, int algo, bool seq_scan, double *out1, double *out2)
{
    const int N = cv.N, tid = threadIdx.x, nt = blockDim.x;
    const double v = th[0], tau = th[1], ton = th[2], b1 = th[3], b2 = th[4], A = th[5], R = th[6];

    // (a) per-step loading increments rho_i*delta_i; zero before onset (the reference `continue`s,
    //     which leaves the counter unchanged)               ConstantElongationSim.m:33-36,57-60
    for (int i = tid; i < N - 1; i += nt) {
        double r = R + th[7 + i];                             // SumofSquares...m:45
        r = r < 0.0 ? 0.0 : r;
        w.rd[i] = (cv.tg[i] < ton) ? 0.0 : __dmul_rn(r, cv.dtg[i]);
    }
    for (int j = tid; j < N; j += nt) { w.F1[j] = 0.0; w.F2[j] = 0.0; }
    __syncthreads();

    // (b)+(c) warp 0 scans the loading counter while warps 1 and 2 build the response tables of the
    //     two colours; then every thread sums its time points.  One loop set at a time: the basal
    //     clamp sits inside the per-set loop in the reference (GetFluorFromPolPos.m:47,57,69).
    const double tv = tau * v;
    const double L1 = C.L_ms2 + tv, L2 = C.L_pp7 + tv;        // :19-20
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = 0; s < C.nsets; ++s) {
        const double s1 = C.ms2_start[s], e1 = C.ms2_end[s], f1 = C.ms2_loopn[s] / 24.0;
        const double s2 = C.pp7_start[s], e2 = C.pp7_end[s], f2 = C.pp7_loopn[s] / 24.0;
        if (s == 0 && warp == 0) scan_counts(N, w.rd, w.K, seq_scan);
        if (algo == TC_ALGO_TOEPLITZ) {
            // lag thresholds of the piecewise response, by exact predicate on p = v*(d*lag):
            // ramp = [la, le), plateau = [lb, lL)
            if (warp == 1 || warp == 2) {
                const bool c2 = warp == 2;
                const double xs = c2 ? s2 : s1, xe = c2 ? e2 : e1, xL = c2 ? L2 : L1, xf = c2 ? f2 : f1;
                int *thr = w.thr + (c2 ? 4 : 0);
                if (lane < 4) {
                    const double x = lane == 0 ? xs : (lane == 3 ? xL : xe);
                    thr[lane] = first_lag(v, cv.d, x, N, (lane & 1) == 0);   // >s, >=e, >e, >=L
                }
                __syncwarp();
                const int la = thr[0], le = thr[1];
                double *G = c2 ? w.G2 : w.G1;
                const double sc = xf / (xe - xs);
                for (int lag = la + lane; lag < le; lag += 32) G[lag] = (v * (cv.d * (double)lag) - xs) * sc;
            }
            __syncthreads();
            const int la1 = w.thr[0], le1 = w.thr[1], lb1 = w.thr[2], lL1 = w.thr[3];
            const int la2 = w.thr[4], le2 = w.thr[5], lb2 = w.thr[6], lL2 = w.thr[7];
            for (int j = tid; j < N; j += nt) {
                double a1 = 0.0, a2 = 0.0;
                const int h1 = min(le1 - 1, j), h2 = min(le2 - 1, j);
                for (int lag = la1; lag <= h1; ++lag)
                    a1 = fma(w.K[j - lag + 1] - w.K[j - lag], w.G1[lag], a1);
                for (int lag = la2; lag <= h2; ++lag)
                    a2 = fma(w.K[j - lag + 1] - w.K[j - lag], w.G2[lag], a2);
                // plateau: cohorts with lag in [lb, lL) are whole polymerases -> exact count
                if (j >= lb1 && lb1 < lL1) a1 = fma(f1, w.K[j - lb1 + 1] - w.K[max(j - lL1 + 1, 0)], a1);
                if (j >= lb2 && lb2 < lL2) a2 = fma(f2, w.K[j - lb2 + 1] - w.K[max(j - lL2 + 1, 0)], a2);
                double m1 = w.F1[j] + a1, m2 = w.F2[j] + a2;
                w.F1[j] = m1 < b1 ? b1 : m1;                  // :57
                w.F2[j] = m2 < b2 ? b2 : m2;                  // :69
            }
        } else {
            __syncthreads();
            for (int j = tid; j < N; j += nt) {
                double a1 = 0.0, a2 = 0.0;
                const double tj = cv.tg[j];
                for (int i = 0; i < j; ++i) {
                    const double ni = w.K[i + 1] - w.K[i];
                    if (ni > 0.0) {
                        const double p = v * (tj - cv.tg[i]);
                        a1 = fma(ni, loop_response(p, s1, e1, L1, f1), a1);
                        a2 = fma(ni, loop_response(p, s2, e2, L2, f2), a2);
                    }
                }
                double m1 = w.F1[j] + a1, m2 = w.F2[j] + a2;
                w.F1[j] = m1 < b1 ? b1 : m1;
                w.F2[j] = m2 < b2 ? b2 : m2;
            }
        }
        __syncthreads();
    }
    if (out1) for (int j = tid; j < N; j += nt) { out1[j] = A * w.F1[j]; out2[j] = w.F2[j]; }

    // (d) MS2 *= A, interp1 back to the experimental times, NaN-skipping residual sum of squares
    //     SumofSquares...m:51-64
    double acc = 0.0, dummy = 0.0;
    for (int j = tid; j < N; j += nt) {
        const int k = cv.ik[j];
        if (k >= 0) {
            const double wj = cv.iw[j];
            const double m1 = A * w.F1[k], m1n = A * w.F1[k + 1];
            const double r1 = cv.ms2[j] - (m1 + wj * (m1n - m1));
            const double r2 = cv.pp7[j] - (w.F2[k] + wj * (w.F2[k + 1] - w.F2[k]));
            if (r1 == r1) acc += r1 * r1;                      // nansum
            if (r2 == r2) acc += r2 * r2;
        }
    }
    block_sum2(acc, dummy, w.red);
    return acc;
}

}  // namespace tc
