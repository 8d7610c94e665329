"""Development aid: throughput and per-phase cycles of the sampler for a given layout (0 auto, 1 big, 2 chain-per-warp).
   python scripts/warp_perf.py <chains per cell> <nsimu> <burn> <layout> [reps]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
per = int(sys.argv[1]); nsimu = int(sys.argv[2]); burn = int(sys.argv[3]); layout = int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ncell = int(os.environ.get("NCELL", "299"))
cc = np.repeat(np.arange(ncell, dtype=np.int32), per)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn, layout=layout)
for rep in range(reps):
    out = cells.mcmc_run(opts, cc, *inputs)
cnt = out["counters"].sum(axis=0); tot = cc.size * nsimu
pc = out["counters"][:, 8:16].sum(axis=0) / tot
print("layout %d: %d chains x %d steps (burn %d): kernel %.3f s -> %.3e steps/s; evals/step %.2f (executed %.2f) acc %.3f adapt %d cholfail %d" % (
    layout, cc.size, nsimu, burn, out["kernel_seconds"], tot / out["kernel_seconds"], cnt[0] / tot, pc[7], (cnt[1] + cnt[2]) / tot, cnt[4], cnt[5]))
if layout == 2 or (layout == 0 and cc.size >= 6 * 148):
    print("  warp kernel cycles/step: generate %.0f (randomness %.0f) steps %.0f barrier %.0f adapt %.0f (scatter %.0f, +chol %.0f) total %.0f" % (
        pc[0], pc[3], pc[1], pc[2], pc[5], pc[4], pc[6], pc[0] + pc[1] + pc[2] + pc[5]))
else:
    print("  cycles/step: generate %.0f steps %.0f - %.0f %.0f %.0f adapt %.0f total %.0f" % (*pc[:6], pc[:6].sum()))
