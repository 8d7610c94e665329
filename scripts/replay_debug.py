"""Development aid: replay one chain against the C oracle and show the first step where the two differ."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transcriptioncycleinference_b200 import _lib, setup_cell
from transcriptioncycleinference_b200.engine import Cells
from oracle import c_oracle, forward_literal
g = dict(np.load("tests/golden/cells.npz"))
c = int(sys.argv[1]); nsimu = int(sys.argv[2]); burn = int(sys.argv[3]); seed = int(sys.argv[4]); layout = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cons = c_oracle.Construct.from_dict(forward_literal.CONSTRUCTS["P2P-MS2v5-LacZ-PP7v4"])
cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"])
cc = np.array([c], dtype=np.int32)
inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(seed))
N = int(g["N"][c]); npar = 7 + N; o = int(g["off"][c])
r = np.random.default_rng(seed + 1)
st = dict(z1=r.standard_normal((1, nsimu, cells.ld)), u1=r.random((1, nsimu)), z2=r.standard_normal((1, nsimu, cells.ld)), u2=r.random((1, nsimu)), chi2=np.zeros((1, nsimu)))
st["chi2"][0] = r.chisquare(1 + 2 * N, nsimu)
extra = dict(qcovadj_always=int(os.environ.get("QA", "0")))
opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, layout=layout, **extra)
out = cells.mcmc_run(opts, cc, *inputs, replay=st, want_flags=True)
sti = dict(z1=st["z1"][0][:, :npar], u1=st["u1"][0], z2=st["z2"][0][:, :npar], u2=st["u2"][0], chi2=st["chi2"][0])
ref = c_oracle.dram(cons, g["t"][o:o + N], g["ms2"][o:o + N], g["pp7"][o:o + N], c_oracle.default_opts(nsimu, burn, **extra), *[x[0, :npar] for x in inputs], streams=sti)
d = out["flags"][0] != ref["flags"]
print("differing flags:", int(d.sum()), "counters gpu", out["counters"][0][:8], "oracle", ref["counters"])
if d.any():
    k = int(np.argmax(d))
    print("first at step", k, "gpu flag", out["flags"][0][k], "oracle", ref["flags"][k])
    for j in range(max(0, k - 2), k + 2):
        print(j, "ss gpu %.17g oracle %.17g  s2 gpu %.17g oracle %.17g  maxdx %.3g" % (out["sschain"][0][j], ref["sschain"][j], out["s2chain"][0][j], ref["s2chain"][j], np.abs(out["chain"][0][j, :npar] - ref["chain"][j]).max()))
    # the stage-1 / stage-2 proposals of step k from the oracle's previous row
    old = ref["chain"][k - 1]
    R = np.sqrt(inputs[1][0, :npar])
    y1 = old + sti["z1"][k] * R; y2 = old + sti["z2"][k] * R / 5
    for nm, y in (("y1", y1), ("y2", y2)):
        so = c_oracle.ss(cons, g["t"][o:o + N], g["ms2"][o:o + N], g["pp7"][o:o + N], y)
        th = np.zeros((1, cells.ld)); th[0, :npar] = y
        sg = cells.ss_batch(cc, th)[0]; sp = cells.ss_batch(cc, th, algo=0)[0]
        print(nm, "oob", bool(np.any(y < inputs[2][0, :npar]) or np.any(y > inputs[3][0, :npar])), "ss oracle %.17g gpu toeplitz %.17g pairs %.17g" % (so, sg, sp))
else:
    print("max |chain diff| %.3g" % np.abs(out["chain"][0][:, :npar] - ref["chain"]).max())
