import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
nch = int(sys.argv[1]); nsimu = int(sys.argv[2]); burn = int(sys.argv[3])
cc = (np.arange(nch) % 299).astype(np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn)
for rep in range(2):
    out = cells.mcmc_run(opts, cc, *inputs)
cnt = out["counters"].sum(axis=0); tot = nch * nsimu
pc = out["counters"][:, 8:16].sum(axis=0) / tot
print("%d chains x %d steps (burn %d): kernel %.3f s -> %.3e steps/s; evals/step %.2f spec evals/step %.2f acc %.3f" % (nch, nsimu, burn, out["kernel_seconds"], tot / out["kernel_seconds"], cnt[0] / tot, pc[7], (cnt[1] + cnt[2]) / tot))
print("  cycles/step: generate %.0f speculate %.0f emit-rej %.0f accept %.0f s2+state %.0f adapt %.0f total %.0f" % (*pc[:6], pc[:6].sum()))
tot_c = out["counters"][:, 8:14].sum(axis=1) / nsimu
Ns = g["N"][cc]
print("  per-chain cycles/step: mean %.0f max %.0f min %.0f; by N:" % (tot_c.mean(), tot_c.max(), tot_c.min()), {int(n): int(tot_c[Ns == n].mean()) for n in np.unique(Ns)})
acc = (out["counters"][:, 1] + out["counters"][:, 2]) / nsimu
print("  acceptance by chain: mean %.3f min %.3f max %.3f; corr(cycles, acc) %.2f" % (acc.mean(), acc.min(), acc.max(), np.corrcoef(tot_c, acc)[0, 1]))
i = int(np.argmax(tot_c)); j = int(np.argmin(tot_c))
for tag, q in (("slowest", i), ("fastest", j)):
    c = out["counters"][q]
    nbatch_est = None
    print("  %s chain %d (N=%d, acc %.3f, evals/step %.2f, spec evals/step %.2f): generate %.0f speculate %.0f emit-rej %.0f accept %.0f s2+state %.0f adapt %.0f" % (tag, q, Ns[q], acc[q], c[0] / nsimu, c[15] / nsimu, *(c[8:14] / nsimu)))
