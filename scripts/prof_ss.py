"""Small driver for ncu: a few launches of the batched SS kernel (and optionally a short DRAM run)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells

mode = sys.argv[1] if len(sys.argv) > 1 else "ss"
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
rng = np.random.default_rng(0)
if mode == "ss":
    n = 299 * 256
    cid = np.repeat(np.arange(299, dtype=np.int32), n // 299)
    th = np.zeros((cid.size, cells.ld))
    for c in range(299):
        N = int(g["N"][c]); m = cid == c
        lo = np.concatenate([[0.5, 0, 0, 0, 0, 0, 5], -8 * np.ones(N)]); hi = np.concatenate([[4, 6, 6, 3, 3, 1, 25], 8 * np.ones(N)])
        th[m, :7 + N] = lo + (hi - lo) * rng.random((m.sum(), 7 + N))
    d_th = torch.from_numpy(th).cuda(); d_cid = torch.from_numpy(cid).cuda(); d_out = torch.zeros(cid.size, dtype=torch.float64, device="cuda")
    algo = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for rep in range(3):
        cells.ss_batch_device(cid.size, d_cid.data_ptr(), d_th.data_ptr(), cells.ld, d_out.data_ptr(), algo=algo)
    torch.cuda.synchronize()
    print("ss ok", float(d_out.sum()))
else:
    nsimu = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    cc = np.tile(np.arange(299, dtype=np.int32), 2)
    inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
    opts = _lib.default_opts(nsimu=nsimu, burnintime=nsimu // 2, n_burn=nsimu // 2)
    out = cells.mcmc_run(opts, cc, *inputs)
    print("mcmc ok", out["kernel_seconds"])
