"""Debug: GPU production run vs GPU replay of the dumped Philox streams vs the C / NumPy oracles on the same streams."""
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
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 20201028
uidmul = int(sys.argv[5]) if len(sys.argv) > 5 else (1 << 20)
cc_all = np.arange(299, dtype=np.int32)
inputs_all = setup_cell.chain_inputs(cells, cc_all, np.random.default_rng(1000))
uid_all = cc_all.astype(np.uint64) * np.uint64(uidmul)
cc = cc_all[chains]; inputs = [x[chains] for x in inputs_all]; uid = uid_all[chains]
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed)
prod = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid, want_flags=True)
nch = len(cc)
st = dict(z1=np.zeros((nch, nsimu, cells.ld)), z2=np.zeros((nch, nsimu, cells.ld)), u1=np.zeros((nch, nsimu)),
          u2=np.zeros((nch, nsimu)), chi2=np.zeros((nch, nsimu)))
dumps = []
for i, c in enumerate(cc):
    N = int(g["N"][c]); npar = 7 + N
    d = _lib.rng_dump(seed, int(uid[i]), npar, 1 + 2 * N, nsimu)
    dumps.append(d)
    st["z1"][i, :, :npar] = d["z1"]; st["z2"][i, :, :npar] = d["z2"]
    st["u1"][i] = d["u1"]; st["u2"][i] = d["u2"]; st["chi2"][i] = d["chi2"]
ropts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed, replay=1)
rep = cells.mcmc_run(ropts, cc, *inputs, chain_uid=uid, replay=st, want_flags=True)


def first_bad(a, b):
    same = a == b
    return -1 if same.all() else int(np.argmin(same))


for i, c in enumerate(cc):
    N = int(g["N"][c]); o = int(g["off"][c]); npar = 7 + N
    r = c_oracle.dram(cons, g["t"][o:o + N], g["ms2"][o:o + N], g["pp7"][o:o + N], c_oracle.default_opts(nsimu, burn),
                      *[x[i, :npar] for x in inputs], streams=dumps[i])
    print("chain %d (N=%d): first flag mismatch prod-vs-replay %d, prod-vs-oracle %d, replay-vs-oracle %d; acc prod %.4f replay %.4f oracle %.4f" % (
        c, N, first_bad(prod["flags"][i], rep["flags"][i]), first_bad(prod["flags"][i], r["flags"]), first_bad(rep["flags"][i], r["flags"]),
        (prod["flags"][i] & 1).mean(), (rep["flags"][i] & 1).mean(), (r["flags"] & 1).mean()))
    print("    max|chain diff| prod-oracle %.3e replay-oracle %.3e; counters prod %s" % (
        np.abs(prod["chain"][i][:, :npar] - r["chain"]).max(), np.abs(rep["chain"][i][:, :npar] - r["chain"]).max(), prod["counters"][i][:8]))
    k = first_bad(prod["flags"][i], r["flags"])
    if k >= 0:
        print("    step %d: flags prod %d replay %d oracle %d; ss prod %.10g replay %.10g oracle %.10g" % (
            k, prod["flags"][i][k], rep["flags"][i][k], r["flags"][k], prod["sschain"][i][k], rep["sschain"][i][k], r["sschain"][k]))
        # proposal actually made at step k if it was accepted: row k - row k-1
        for nm, ch in (("prod", prod["chain"][i][:, :npar]), ("replay", rep["chain"][i][:, :npar]), ("oracle", r["chain"])):
            dlt = ch[k] - ch[k - 1]
            print("      %s: |row k - row k-1| = %.4g (max comp %.4g)" % (nm, np.linalg.norm(dlt), np.abs(dlt).max()))
