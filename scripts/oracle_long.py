"""Long-run exactness: GPU production run vs the C oracle fed with the device's own Philox streams."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
from oracle import c_oracle, forward_literal
g = dict(np.load("tests/golden/cells.npz"))
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
cons = c_oracle.Construct.from_dict(forward_literal.CONSTRUCTS["P2P-MS2v5-LacZ-PP7v4"])
nsimu, burn = int(sys.argv[1]), int(sys.argv[2])
chains = [int(x) for x in sys.argv[3].split(",")]
cc_all = np.arange(299, dtype=np.int32)
inputs_all = setup_cell.chain_inputs(cells, cc_all, np.random.default_rng(1000))
uid_all = cc_all.astype(np.uint64) * np.uint64(1 << 20)
seed = 20201028
cc = cc_all[chains]; inputs = [x[chains] for x in inputs_all]; uid = uid_all[chains]
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed)
out = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid, want_flags=True)
for i, c in enumerate(cc):
    N = int(g["N"][c]); o = int(g["off"][c]); npar = 7 + N
    d = _lib.rng_dump(seed, int(uid[i]), npar, 1 + 2 * N, nsimu)
    r = c_oracle.dram(cons, g["t"][o:o + N], g["ms2"][o:o + N], g["pp7"][o:o + N], c_oracle.default_opts(nsimu, burn),
                      *[x[i, :npar] for x in inputs], streams=d)
    same = out["flags"][i] == r["flags"]
    first_bad = int(np.argmin(same)) if not same.all() else -1
    print("chain %d: flags identical %s (first mismatch at step %d), max |chain diff| %.3e, gpu acc %.4f oracle acc %.4f, gpu mean v %.4f oracle mean v %.4f" % (
        c, same.all(), first_bad, np.abs(out["chain"][i][:, :npar] - r["chain"]).max(), (out["flags"][i] & 1).mean(), (r["flags"] & 1).mean(),
        out["chain"][i][:, 0].mean(), r["chain"][:, 0].mean()))
    if first_bad >= 0:
        k = first_bad
        print("   step %d: gpu flag %d oracle flag %d; ss gpu %.10g oracle %.10g (prev step ss %.10g / %.10g); s2 %.6g / %.6g" % (
            k, out["flags"][i][k], r["flags"][k], out["sschain"][i][k], r["sschain"][k], out["sschain"][i][k-1], r["sschain"][k-1], out["s2chain"][i][k-1], r["s2chain"][k-1]))
