import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nch = int(sys.argv[1]); nsimu = int(sys.argv[2])
cc = (np.arange(nch) % 299).astype(np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=nsimu // 2, n_burn=nsimu // 2)
out = cells.mcmc_run(opts, cc, *inputs)
print("ok", out["kernel_seconds"], out["counters"][:, 0].sum())
