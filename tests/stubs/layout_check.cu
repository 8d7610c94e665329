// Host-side check of the shared-memory layouts the kernels rely on (compiled and run by tests/test_layout_invariants.py;
// needs nvcc, no GPU).  It includes the product's translation unit, so the formulas checked are the ones the kernels use.
#include <cstdio>
#include <set>
#include "../../transcriptioncycleinference_b200/csrc/tc_mcmc.cu"

static int fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (++fails < 20) { std::printf("FAIL: " __VA_ARGS__); std::printf("\n"); } } } while (0)

int main()
{
    const size_t optin = 232448;                      // 227 KB: B200's opt-in shared memory per block
    for (int N = 3; N <= 460; ++N) {
        // (1) the permuted model grid: a bijection of [0, N) into [0, (N + 3) & ~3), and the four loads of lane l in the block
        //     starting at r0 (steps r0 + 4 l + e) sit at r0 + e q + l
        const int N4 = (N + 3) & ~3;
        std::set<int> seen;
        for (int i = 0; i < N; ++i) {
            const int p = cell_perm(N, i);
            CHECK(p >= 0 && p < N4, "cell_perm(%d, %d) = %d outside [0, %d)", N, i, p, N4);
            CHECK(seen.insert(p).second, "cell_perm(%d, %d) = %d twice", N, i, p);
            const int r0 = i & ~127, rem = N - r0, q = ((rem < 128 ? rem : 128) + 3) >> 2, lane = (i & 127) >> 2, e = i & 3;
            CHECK(p == r0 + e * q + lane, "cell_perm(%d, %d) != block + e q + lane", N, i);
        }
        // (2) the forward-model scratch: thr (4) + alignment (1) + K, S as 32-bit integers with 3 leading pad ints and room for
        //     the 16-byte store of the last lane (index n + 3 = N + 2) + n, F1, F2 of N + 2 doubles
        const int ks = work_ks_doubles(N);
        CHECK(ks % 2 == 0 && 2 * ks >= 3 + (N + 3), "K/S area of %d doubles too small for N = %d", ks, N);
        CHECK(tc::work_doubles(N) >= 4 + 1 + 2 * ks + 3 * (N + 2), "work_doubles(%d)", N);
        CHECK(tc::cell_doubles(N) >= 2 * N4 + 3 * (N + 1) + (N + 2) / 2 + 1, "cell_doubles(%d)", N);
        // (3) a ring slot: two increment vectors of s1 >= npar doubles at odd offsets + 8 scalars, even size; Z (2 npar doubles
        //     from the slot base, big layout) must not reach the scalars
        const int s1 = dram_s1(N), slot = dram_slot(N), npar = 7 + N;
        CHECK(s1 % 2 == 0 && s1 >= npar && slot % 2 == 0 && slot == 2 * s1 + 10, "slot of N = %d", N);
        CHECK(2 * npar <= slot - 8, "Z overlaps the scalars at N = %d", N);
        // (4) budgets: what the kernel carves (cell, <= 3 alignment doubles, vectors, ring, per-warp areas) fits what the host asks for
        for (int big = 0; big < 2; ++big) {
            const int wsz = dram_wsz(N, big);
            CHECK(wsz % 2 == 0 && wsz >= ((tc::work_doubles(N) + 1) & ~1), "wsz(%d, %d)", N, big);
            const long long carved = tc::cell_doubles(N) + 3 + (long long)(big ? 7 : 10) * npar + (long long)(big ? RING / 2 : RING) * slot + (long long)SPEC * wsz;
            CHECK(carved <= dram_smem_doubles(N, big), "carve of N = %d (big %d) exceeds its budget", N, big);
            if (!big) CHECK((long long)(big ? RING / 2 : RING) * slot + (long long)SPEC * wsz >= chol_ws_doubles(npar) || sizeof(double) * (size_t)dram_smem_doubles(N, 0) > optin,
                            "Cholesky workspace does not fit ring + per-warp areas at N = %d", N);
        }
        // (5) chain-per-warp region: two slots of ldp >= npar at odd offsets, then the overlay (cell + scratch | normals | panel)
        CHECK(wk_ldp(N) % 2 == 0 && wk_ldp(N) >= npar && wk_region(N) % 2 == 0, "warp region of N = %d", N);
        CHECK(wk_overlay(N) >= wk_cell_sz(N) + ((tc::work_doubles(N) + 3) & ~1), "warp overlay of N = %d", N);
    }
    // the series lengths each layout takes on a B200 (static shared memory of the kernels: < 2.5 KB / 13 KB)
    int max_reg = 0, max_big = 0, max_warp = 0;
    for (int N = 3; N <= 460; ++N) {
        if (sizeof(double) * (size_t)dram_smem_doubles(N, 0) + 2560 <= optin) max_reg = N;
        if (sizeof(double) * (size_t)dram_smem_doubles(N, 1) + 2560 <= optin && 7 + N <= 8 * SPEC * GENB_MAXT) max_big = N;
        if (sizeof(double) * (size_t)wk_region(N) * WK_WARPS + 13312 <= optin) max_warp = N;
    }
    std::printf("max N: regular %d, big %d, chain-per-warp %d\n", max_reg, max_big, max_warp);
    CHECK(2 * (sizeof(double) * (size_t)dram_smem_doubles(129, 0) + 2560 + 1024) <= optin + 1024, "two CTAs per SM at N = 129");
    CHECK(max_big == 441, "big layout limit %d (include/tcmcmc.h says 441)", max_big);
    std::printf(fails ? "layout check: %d failure(s)\n" : "layout check ok\n", fails);
    return fails ? 1 : 0;
}
