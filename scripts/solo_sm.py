"""Development aid: how much faster is a chain of the CTA-per-chain sampler when its CTA has the SM to itself?
The same 148 chains are fitted (a) alone: 148 CTAs, one per SM, (b) among 296 chains: two CTAs per SM."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
cc = np.arange(296, dtype=np.int32)
uid = cc.astype(np.uint64)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
opts.layout = _lib.LAYOUT_CTA if hasattr(_lib, "LAYOUT_CTA") else opts.layout
res = {}
for tag, sel in (("pair", np.arange(296)), ("solo", np.arange(0, 296, 2)), ("pair", np.arange(296)), ("solo", np.arange(0, 296, 2))):
    out = cells.mcmc_run(opts, cc[sel], *[x[sel] for x in inputs], chain_uid=uid[sel])
    cyc = out["counters"][:, 8:14].sum(axis=1) / nsimu
    res[tag] = dict(zip(sel.tolist(), cyc))
    ph = out["counters"][:, 8:14].sum(axis=0) / (len(sel) * nsimu)
    print("%s: %d chains, kernel %.3f s, cycles/step mean %.0f max %.0f; phases gen %.0f spec %.0f commit %.0f . . adapt %.0f" % (
        tag, len(sel), out["kernel_seconds"], cyc.mean(), cyc.max(), ph[0], ph[1], ph[2], ph[5]))
common = sorted(res["solo"].keys())
ratio = np.array([res["solo"][c] / res["pair"][c] for c in common])
print("solo / pair cycles per step: mean %.3f min %.3f max %.3f" % (ratio.mean(), ratio.min(), ratio.max()))
