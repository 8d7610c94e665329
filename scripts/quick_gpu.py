"""First-contact GPU script: FP64 peak, SS batch timing, short MCMC timing (development aid)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells

g = dict(np.load("tests/golden/cells.npz"))
print(_lib.device_info(0))
peak, clk = _lib.measure_fp64_peak(0)
print("DFMA peak %.3e lane-ops/s  (%.1f TFLOP/s)  SM clock %.0f MHz" % (peak, 2 * peak / 1e12, clk))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 299 * 256
cid = np.repeat(np.arange(299, dtype=np.int32), n // 299)
th = np.zeros((cid.size, cells.ld))
for c in range(299):
    N = int(g["N"][c]); m = cid == c
    lo = np.concatenate([[0.5, 0, 0, 0, 0, 0, 5], -8 * np.ones(N)]); hi = np.concatenate([[4, 6, 6, 3, 3, 1, 25], 8 * np.ones(N)])
    th[m, :7 + N] = lo + (hi - lo) * rng.random((m.sum(), 7 + N))
for algo in (1, 0):
    cells.ss_batch(cid[:1000], th[:1000], algo=algo)
    t0 = time.time(); ss = cells.ss_batch(cid, th, algo=algo); dt = time.time() - t0
    print("algo %d: %d evals in %.3f s (host buffers, incl. copies) -> %.3e evals/s" % (algo, cid.size, dt, cid.size / dt))
import torch
d_th = torch.from_numpy(th).cuda(); d_cid = torch.from_numpy(cid).cuda(); d_out = torch.zeros(cid.size, dtype=torch.float64, device="cuda")
for algo in (1, 0):
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); cells.ss_batch_device(cid.size, d_cid.data_ptr(), d_th.data_ptr(), cells.ld, d_out.data_ptr(), algo=algo); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("algo %d device-resident: %d evals in %.3f ms -> %.3e evals/s; W_op frac of DFMA peak %.3f" % (algo, cid.size, ms, cid.size / ms * 1e3, cid.size / ms * 1e3 * 80940 / peak))
    assert np.allclose(d_out.cpu().numpy(), ss, rtol=1e-9)
for nsimu, burn, nrep in ((2000, 1000, 1), (2000, 1000, 8)):
    cc = np.tile(np.arange(299, dtype=np.int32), nrep)
    inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
    t0 = time.time(); out = cells.mcmc_run(opts, cc, *inputs); dt = time.time() - t0
    cnt = out["counters"].sum(axis=0)
    print("mcmc %d chains x %d steps: wall %.3f s kernel %.3f s -> %.3e steps/s, %.3e ss evals/s; evals/step %.2f acc1 %.3f acc2 %.3f oob %.3f adapt %d cholfail %d"
          % (cc.size, nsimu, dt, out["kernel_seconds"], cc.size * nsimu / out["kernel_seconds"], cnt[0] / out["kernel_seconds"],
             cnt[0] / (cc.size * nsimu), cnt[1] / (cc.size * nsimu), cnt[2] / (cc.size * nsimu), cnt[3] / (cc.size * nsimu), cnt[4], cnt[5]))
    pc = out["counters"][:, 8:16].sum(axis=0) / (cc.size * nsimu)
    print("  cycles/step by phase: generate %.0f speculate %.0f resolve[emit-rej %.0f accept %.0f s2+state %.0f] adapt %.0f total %.0f; speculative evals/step %.2f" % (*pc[:6], pc[:6].sum(), pc[7]))
    print("  mean v %.3f tau %.3f ton %.3f sigma %.3f" % (out["mean"][:, 0].mean(), out["mean"][:, 1].mean(), out["mean"][:, 2].mean(), out["sig"][:, 0].mean()))
