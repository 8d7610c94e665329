"""Development aid: the same chains through dram_kernel (layout 0) and dram_warp_kernel (layout 2)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu = int(sys.argv[1]); burn = int(sys.argv[2]); qa = int(sys.argv[3])
cc = np.arange(0, 296, 8, dtype=np.int32)
uid = cc.astype(np.uint64) * np.uint64(1 << 20) + np.uint64(3)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(101))
res = {}
for layout in (2, 0, 1):
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=20201028, layout=layout, qcovadj_always=qa)
    res[layout] = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid, want_flags=True)
for a, b in ((2, 0), (1, 0), (2, 1)):
    d = res[a]["flags"] != res[b]["flags"]
    bad = np.where(d.any(axis=1))[0]
    print("layouts %d vs %d: chains with differing flags:" % (a, b), bad.tolist(), "first steps", [int(np.argmax(d[i])) for i in bad])
    for i in bad[:3]:
        k = int(np.argmax(d[i]))
        print("  chain %d step %d: ss %.15g vs %.15g, max|dx| before %.3g, cholfail %s %s" % (i, k, res[a]["sschain"][i][k], res[b]["sschain"][i][k],
              np.abs(res[a]["chain"][i][k - 1] - res[b]["chain"][i][k - 1]).max(), res[a]["counters"][i][5], res[b]["counters"][i][5]))
