"""Sub-phase cycle breakdown of chain 0 (development aid).  Build the profiling variant first:
   TC_LIBTCMCMC=$PWD/transcriptioncycleinference_b200/libtcmcmc_prof.so TC_NVCC_EXTRA=-DTC_SUBPROF python -m transcriptioncycleinference_b200.build --force
and run this script with the same TC_LIBTCMCMC."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nch = int(sys.argv[1]); nsimu = int(sys.argv[2]); burn = int(sys.argv[3])
first = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cc = ((np.arange(nch) + first) % 299).astype(np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
out = cells.mcmc_run(opts, cc, *inputs)
_lib.debug_subprof()
out = cells.mcmc_run(opts, cc, *inputs)
sp = _lib.debug_subprof().astype(float)
c = out["counters"][0]
rounds, commits = sp[24], sp[25]
print("chain 0 (cell %d, N=%d): %d steps, %d rounds (%.2f steps/round), acc %.3f, kernel %.3f s" % (cc[0], g["N"][cc[0]], nsimu, rounds, commits / max(rounds, 1), (c[1] + c[2]) / nsimu, out["kernel_seconds"]))
print("  phase cycles/step: generate %.0f speculate %.0f commit %.0f - %.0f - %.0f adapt %.0f" % tuple(c[8:14] / nsimu))
names = {0: "gen: randomness + sync", 1: "gen: norms", 2: "gen: dmma (B from L2)", 3: "gen: sync + write + sync (or diag scale)",
         16: "round: A bounds+prior (warp 0)", 17: "round: barrier 1", 18: "round: B tasks + C evaluation (warp 0)", 19: "round: barrier 2", 20: "round: D resolve", 21: "round: flush_run (accepts)", 22: "round: rows + state (warp 0)", 23: "round: barrier 3",
         13: "chol: panel solve (thread 0)", 14: "chol: barrier after solve", 6: "chol: next-diag tile update (warp 0)", 7: "chol: next-diag factorisation (warp 0)", 15: "chol: rest until barrier", 5: "chol: barrier after trailing",
         8: "adapt: means + scatter accumulate", 9: "adapt: M2 rmw + cmean", 10: "adapt: load cov / burn-in scale", 11: "adapt: cholesky", 12: "adapt: write R"}
gcalls = max(sp[26], 1)
print("  generate calls %d (%.1f new steps/call)" % (gcalls, sp[27] / gcalls))
for i, nm in names.items():
    per = gcalls if i < 8 else (rounds if i >= 16 else nsimu / 100)
    print("  %-40s %10.0f cycles/call  %8.0f cycles/step" % (nm, sp[i] / max(per, 1), sp[i] / nsimu))
