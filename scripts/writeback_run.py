"""Config 2 at the reference's code defaults with raw chains stored (two launches): the launch ncu captures for the HBM
write rate of chain write-back (scripts/profile_gpu.sh)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
cc = np.arange(299, dtype=np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1000))
opts = _lib.default_opts(nsimu=20000, burnintime=10000, n_burn=10000, store_chain=1, seed=20201028)
chain = _lib.pinned_empty((299, 10001, cells.ld)); s2 = _lib.pinned_empty((299, 20000))
for _ in range(2):
    out = cells.mcmc_run(opts, cc, *inputs, chain_out=chain, s2chain_out=s2)
    print("kernel %.4f s, drain %.4f s, %.2f GB" % (out["kernel_seconds"], out["drain_seconds"], (chain.nbytes + s2.nbytes) / 1e9))
