"""Development aid: config 2 (299 chains x nsimu steps) under different settings of the solo-SM scheduler
(TC_SOLO_LAG = slices of lag that earn a chain its SM, 0 = off; TC_NSEG = time slices per chain)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
cc = np.arange(299, dtype=np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
ref = None
for lag, nseg in [(0, 32), (0, 128), (1, 128), (2, 128), (4, 128), (2, 64), (2, 256), (3, 256)]:
    os.environ["TC_SOLO_LAG"] = str(lag); os.environ["TC_NSEG"] = str(nseg)
    best = 1e9
    for rep in range(2):
        out = cells.mcmc_run(opts, cc, *inputs)
        best = min(best, out["kernel_seconds"])
    same = "" if ref is None else (" identical" if np.array_equal(ref, out["mean"]) else " DIFFERENT")
    if ref is None: ref = out["mean"].copy()
    cyc = out["counters"][:, 8:14].sum(axis=1) / nsimu
    print("lag %d nseg %3d: kernel %.4f s -> %.2f M chain-steps/s (cycles/step mean %.0f max %.0f)%s" % (
        lag, nseg, best, 299 * nsimu / best / 1e6, cyc.mean(), cyc.max(), same), flush=True)
