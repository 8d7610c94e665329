import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu, burn = int(sys.argv[1]), int(sys.argv[2])
cc = np.arange(299, dtype=np.int32)
rep = np.zeros(299, dtype=np.uint64)
uid = cc.astype(np.uint64) * np.uint64(1 << 20) + rep
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1000))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn, seed=20201028)
out = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
c = out["counters"]
print("kernel %.3f s; acc %.4f evals/step %.4f adapt %d (expected %d) cholfail %d oob/step %.3f dr/step %.3f" % (
    out["kernel_seconds"], (c[:, 1] + c[:, 2]).sum() / (299 * nsimu), c[:, 0].sum() / (299 * nsimu), c[:, 4].sum(),
    299 * ((nsimu // 100) - (burn // 100) + 1), c[:, 5].sum(), c[:, 3].sum() / (299 * nsimu), c[:, 6].sum() / (299 * nsimu)))
acc = (c[:, 1] + c[:, 2]) / nsimu
print("acc per chain: min %.4f median %.4f max %.4f; chains with acc<0.01: %d" % (acc.min(), np.median(acc), acc.max(), (acc < 0.01).sum()))
print("mean v %.3f tau %.3f sigma %.3f" % (out["mean"][:, 0].mean(), out["mean"][:, 1].mean(), out["sig"][:, 0].mean()))
