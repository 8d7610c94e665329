"""Which chains are slow, and why (development aid): per-chain cycles/step with the bench's own inputs."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu, burn, seed_in = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cc = np.arange(299, dtype=np.int32)
uid = cc.astype(np.uint64) * np.uint64(1 << 20)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(seed_in))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn, seed=20201028)
out = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
c = out["counters"]
cyc = c[:, 8:14].sum(axis=1) / nsimu
acc = (c[:, 1] + c[:, 2]) / nsimu
print("kernel %.3f s; cycles/step mean %.0f median %.0f max %.0f" % (out["kernel_seconds"], cyc.mean(), np.median(cyc), cyc.max()))
for i in np.argsort(-cyc)[:6]:
    print("  chain %3d N=%d cycles/step %.0f [gen %.0f eval %.0f commit %.0f adapt %.0f] acc %.3f evals/step %.2f mean v %.3f tau %.3f ton %.3f A %.3f R %.2f" % (
        i, g["N"][i], cyc[i], c[i, 8] / nsimu, c[i, 9] / nsimu, c[i, 10] / nsimu, c[i, 13] / nsimu, acc[i], c[i, 0] / nsimu, *out["mean"][i, [0, 1, 2, 5, 6]]))
