"""BASELINE config 5: synthetic cells x N = 400 time points (npar = 407: the big layout) generated from the
likelihood's own model with known parameters, DRAM fit, throughput + parameter recovery (truth within posterior mean +- 3
sigma for v, tau, t_on).

  python scripts/config5.py [ncells] [n_steps] [N] [ngpus] [random|truth] [chains per cell]

  truth      the chains start AT the true parameters: does the posterior sit where it should (likelihood + sampler correct)?
  random     the reference's random starts (src/TranscriptionCycleMCMC.m:193-210).  With several chains per cell the chains of a
             cell are pooled and recovery is reported over the cells whose chains agree (Rhat < 1.1 for v, tau, t_on), next to
             the fraction of cells that pass."""
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
nch = int(sys.argv[6]) if len(sys.argv) > 6 else 1
t0 = time.time()
cells, truth = synthetic.make_cells(ncells, N, devices=tuple(range(ngpus)))
t_gen = time.time() - t0
cc = np.repeat(np.arange(ncells, dtype=np.int32), nch)
uid = cc.astype(np.uint64) * np.uint64(1 << 20) + np.tile(np.arange(nch, dtype=np.uint64), ncells)
t0 = time.time()
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(5))
nprobe = min(ncells, 2000)
n_obs = np.array([np.sum(~np.isnan(cells.cell(c)[1])) + np.sum(~np.isnan(cells.cell(c)[2])) for c in range(nprobe)])
rms_truth = float(np.median(np.sqrt(cells.ss_batch(np.arange(nprobe, dtype=np.int32), truth[:nprobe]) / n_obs)))   # ~ noise sigma: the model is consistent
if at_truth:
    inputs = (np.repeat(truth, nch, axis=0).copy(),) + tuple(inputs[1:])
print("generated %d cells in %.1f s, chain inputs in %.1f s" % (ncells, t_gen, time.time() - t0), file=sys.stderr, flush=True)
burn = nsteps // 2
opts = _lib.default_opts(nsimu=nsteps, burnintime=burn, n_burn=burn, ngpus=ngpus)
for rep in range(1 if cc.size >= 20000 else 2):          # (a second, warm repetition for the small runs)
    t0 = time.time()
    out = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
    wall = time.time() - t0
cnt = out["counters"]
ks = out["kernel_seconds"]
res = dict(config="config5-scale: %d cells x N=%d x %d chain(s), n_steps=%d, %d GPU(s), %s" % (
               ncells, N, nch, nsteps, ngpus, "chains started AT the truth" if at_truth else "random starts"),
           median_rms_residual_at_truth=rms_truth,
           chain_steps_per_s=cc.size * nsteps / ks, kernel_s=ks, wall_s=wall, gen_s=t_gen,
           ss_evals_per_step=float(cnt[:, 0].sum()) / (cc.size * nsteps),
           accept_rate=float(cnt[:, 1:3].sum()) / (cc.size * nsteps),
           adaptations=int(cnt[:, 4].sum()), chol_fail=int(cnt[:, 5].sum()),
           median_posterior_sigma=float(np.median(out["sig"][:, 0])))
if nch > 1:
    pm, ps, rh = synthetic.pool_chains(out["mean"], out["std"], nch, nsteps - burn + 1)
    ok = np.nanmax(rh, axis=1) < 1.1
    res.update(recovery_v_tau_ton_3sigma_all_cells=synthetic.recovery(truth, pm, ps),
               recovery_v_tau_ton_3sigma_rhat_ok=synthetic.recovery(truth, pm, ps, keep=ok),
               fraction_cells_rhat_below_1p1=float(ok.mean()), median_rhat_v_tau_ton=np.nanmedian(rh, axis=0).tolist())
    mean, std = pm, ps
else:
    mean, std = out["mean"], out["std"]
    res.update(recovery_v_tau_ton_3sigma=synthetic.recovery(truth, mean, std))
res.update(median_post_std_v_tau_ton=[float(np.median(std[:, i])) for i in range(3)],
           median_abs_err_v_tau_ton=[float(np.median(np.abs(truth[:, i] - mean[:, i]))) for i in range(3)])
pc = cnt[:, 8:14].sum(axis=0).astype(float)
res["cycles_per_step"] = dict(zip(["generate", "rounds", "commit", "p3", "p4", "adapt"], (pc / (cc.size * nsteps)).round(0).tolist()))
print(json.dumps(res))
