"""Sub-phase cycle breakdown for the big layout on the synthetic config-5 workload (development aid; see subprof.py for the
build of the profiling variant).  usage: python scripts/subprof5.py [ncells] [n_steps] [N]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell, synthetic
ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
nsimu = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 400
cells, truth = synthetic.make_cells(ncells, N)
cc = np.arange(ncells, dtype=np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(5))
opts = _lib.default_opts(nsimu=nsimu, burnintime=nsimu // 2, n_burn=nsimu // 2)
out = cells.mcmc_run(opts, cc, *inputs)
_lib.debug_subprof()
out = cells.mcmc_run(opts, cc, *inputs)
sp = _lib.debug_subprof().astype(float)
c = out["counters"]
print("kernel %.3f s, %.3e steps/s" % (out["kernel_seconds"], ncells * nsimu / out["kernel_seconds"]))
print("phase cycles/step (all chains): generate %.0f rounds %.0f commit %.0f - %.0f - %.0f adapt %.0f" % tuple(c[:, 8:14].sum(axis=0) / (ncells * nsimu)))
names = {0: "gen: randomness + sync", 1: "gen: norms", 2: "gen: increments (MMA)", 3: "gen: final sync",
         28: "gen tma: issue (thread 0)", 29: "gen tma: packing rule + acquire (wait for data)", 30: "gen tma: rows (LDS + MMA)", 31: "gen tma: release",
         8: "adapt: means + scatter update", 9: "adapt: cmean", 11: "adapt: cholesky (chol_global)", 12: "adapt: write R",
         13: "cholg: accumulate (block 0, all its chains)", 14: "cholg: S build + barrier", 7: "cholg: diagonal block", 15: "cholg: prefetch + solve + write"}
for i, nm in names.items():
    print("  %-48s %12.0f cycles total (chain 0 / block 0)  %8.0f per step of chain 0" % (nm, sp[i], sp[i] / nsimu))
