import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nsimu, burn = int(sys.argv[1]), int(sys.argv[2])
cc = np.arange(299, dtype=np.int32); uid = np.arange(299, dtype=np.uint64)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
a = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
b = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
print("run-to-run identical:", np.array_equal(a["mean"], b["mean"]), np.array_equal(a["counters"][:, :8], b["counters"][:, :8]))
sub = np.array([0, 7, 150, 298])
c = cells.mcmc_run(opts, cc[sub], *[x[sub] for x in inputs], chain_uid=uid[sub])
print("subset identical:", np.array_equal(a["mean"][sub], c["mean"]), np.abs(a["mean"][sub] - c["mean"]).max())
print("acc rate full", (a["counters"][:, 1] + a["counters"][:, 2]).sum() / (299 * nsimu), "subset", (c["counters"][:, 1] + c["counters"][:, 2]).sum() / (4 * nsimu), (a["counters"][sub, 1] + a["counters"][sub, 2]).sum() / (4 * nsimu))
bad = np.flatnonzero(np.any(a["mean"] != b["mean"], axis=1)); print("chains differing run-to-run:", bad[:20], len(bad))
