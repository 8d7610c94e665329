"""BASELINE config 5 at a single-GPU scale: synthetic cells x N = 400 time points (npar = 407: the big layout), DRAM fit
from random starts, throughput + parameter recovery.  usage: python scripts/config5.py [ncells] [n_steps] [N] [ngpus] [truth]   (truth: start the chains at the true parameters)"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from transcriptioncycleinference_b200 import _lib, setup_cell, synthetic  # noqa: E402

ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
N = int(sys.argv[3]) if len(sys.argv) > 3 else 400
ngpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
at_truth = len(sys.argv) > 5 and sys.argv[5] == "truth"
t0 = time.time()
cells, truth = synthetic.make_cells(ncells, N, devices=tuple(range(ngpus)))
t_gen = time.time() - t0
cc = np.arange(ncells, dtype=np.int32)
t0 = time.time()
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(5))
n_obs = np.array([np.sum(~np.isnan(cells.cell(c)[1])) + np.sum(~np.isnan(cells.cell(c)[2])) for c in range(min(ncells, 2000))])
rms_truth = float(np.median(np.sqrt(cells.ss_batch(cc[:n_obs.size], truth[:n_obs.size]) / n_obs)))   # ~ noise sigma if the model is consistent
if at_truth:
    inputs = (truth.copy(),) + tuple(inputs[1:])
print("generated %d cells in %.1f s, chain inputs in %.1f s" % (ncells, t_gen, time.time() - t0), file=sys.stderr, flush=True)
opts = _lib.default_opts(nsimu=nsteps, burnintime=nsteps // 2, n_burn=nsteps // 2, ngpus=ngpus)
for rep in range(1 if ncells >= 20000 else 2):          # (a second, warm repetition for the small runs)
    t0 = time.time()
    out = cells.mcmc_run(opts, cc, *inputs)
    wall = time.time() - t0
cnt = out["counters"]
ks = out["kernel_seconds"]
rec = synthetic.recovery(truth, out["mean"], out["std"])
pc = cnt[:, 8:14].sum(axis=0).astype(float)
print(json.dumps(dict(config="config5-scale: %d cells x N=%d, n_steps=%d, %d GPU(s)%s" % (ncells, N, nsteps, ngpus, ", chains started AT the truth" if at_truth else ""),
                      median_rms_residual_at_truth=rms_truth,
                      chain_steps_per_s=ncells * nsteps / ks, kernel_s=ks, wall_s=wall, gen_s=t_gen,
                      ss_evals_per_step=float(cnt[:, 0].sum()) / (ncells * nsteps),
                      accept_rate=float(cnt[:, 1:3].sum()) / (ncells * nsteps),
                      adaptations=int(cnt[:, 4].sum()), chol_fail=int(cnt[:, 5].sum()),
                      recovery_v_tau_ton_3sigma=rec, median_posterior_sigma=float(np.median(out["sig"][:, 0])),
                      median_post_std_v_tau_ton=[float(np.median(out["std"][:, i])) for i in range(3)],
                      median_abs_err_v_tau_ton=[float(np.median(np.abs(truth[:, i] - out["mean"][:, i]))) for i in range(3)],
                      cycles_per_step=dict(zip(["generate", "rounds", "commit", "p3", "p4", "adapt"],
                                               (pc / (ncells * nsteps)).round(0).tolist())))))
