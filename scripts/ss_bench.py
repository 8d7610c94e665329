"""Throughput of the batched ssfun kernel alone, device-resident inputs (development aid)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
rng = np.random.default_rng(0); per = 1024
cid = np.repeat(np.arange(299, dtype=np.int32), per)
th = np.zeros((cid.size, cells.ld))
for c in range(299):
    N = int(g["N"][c]); m = slice(c * per, (c + 1) * per)
    lo = np.concatenate([[0.5, 0, 0, 0, 0, 0, 5], -8 * np.ones(N)]); hi = np.concatenate([[4, 6, 6, 3, 3, 1, 25], 8 * np.ones(N)])
    th[m, :7 + N] = lo + (hi - lo) * rng.random((per, 7 + N))
d_th = torch.from_numpy(th).cuda(); d_cid = torch.from_numpy(cid).cuda(); d_out = torch.zeros(cid.size, dtype=torch.float64, device="cuda")
for algo in (1, 0):
    ms = []
    for it in range(8):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); cells.ss_batch_device(cid.size, d_cid.data_ptr(), d_th.data_ptr(), cells.ld, d_out.data_ptr(), algo=algo); e1.record(); torch.cuda.synchronize()
        if it >= 3: ms.append(e0.elapsed_time(e1))
    print("algo %d: %.3f ms -> %.1f M evals/s" % (algo, np.mean(ms), cid.size / np.mean(ms) / 1e3))
