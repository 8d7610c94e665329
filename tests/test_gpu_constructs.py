"""Multi-set constructs on the GPU vs the oracle (per-set basal clamp) and t_start/t_end
windows (truncated cells have their own t_interp)."""
import numpy as np
import pytest

from conftest import random_theta

pytestmark = pytest.mark.gpu


def test_two_set_construct_ss_and_forward(cells_npz, orc):
    from transcriptioncycleinference_b200 import _lib, register_construct
    from transcriptioncycleinference_b200.engine import Cells
    if _lib.device_count() < 1:
        pytest.skip("no CUDA device")
    co, _ = orc
    d = dict(L_MS2=6.0, L_PP7=6.3, MS2_start=[0.024, 2.0], MS2_end=[1.299, 2.6], MS2_loopn=[24.0, 12.0],
             PP7_start=[4.292, 5.8], PP7_end=[5.758, 6.0], PP7_loopn=[24.0, 6.0])
    register_construct("gpu-two-sets", d["L_MS2"], d["L_PP7"], d["MS2_start"], d["MS2_end"], d["MS2_loopn"],
                       d["PP7_start"], d["PP7_end"], d["PP7_loopn"])
    cons2 = co.Construct.from_dict(d)
    sub = np.arange(0, 299, 10)
    N = cells_npz["N"][sub]; off = np.concatenate([[0], np.cumsum(N)]).astype(np.int64)
    t = np.concatenate([cells_npz["t"][cells_npz["off"][c]:cells_npz["off"][c + 1]] for c in sub])
    m2 = np.concatenate([cells_npz["ms2"][cells_npz["off"][c]:cells_npz["off"][c + 1]] for c in sub])
    p7 = np.concatenate([cells_npz["pp7"][cells_npz["off"][c]:cells_npz["off"][c + 1]] for c in sub])
    packed = dict(N=N.astype(np.int32), off=off, t=t, ms2=m2, pp7=p7)
    cells = Cells.from_packed(packed["N"], off, t, m2, p7, construct="gpu-two-sets")
    rng = np.random.default_rng(0)
    n = 4000
    cid = rng.integers(0, len(sub), n).astype(np.int32)
    th = np.zeros((n, cells.ld))
    for i, c in enumerate(cid):
        th[i, :7 + int(N[c])] = random_theta(rng, int(N[c]), i % 2 == 0)
        th[i, 3] = rng.uniform(0, 6); th[i, 4] = rng.uniform(0, 6)     # basal levels that actually clamp
    ref = co.ss_batch(cons2, packed, cid, th)
    for algo in (0, 1):
        got = cells.ss_batch(cid, th, algo=algo)
        rel = np.abs(got - ref) / np.abs(ref)
        assert (rel >= 1e-10).sum() <= 1, (algo, rel.max())
    f1, f2 = cells.forward(cid[:50], th[:50], on_raw_grid=True)
    for i in range(50):
        c = cid[i]; nn = int(N[c]); tt = t[off[c]:off[c] + nn]
        r1, r2 = co.model_on_grid(cons2, th[i, :7 + nn], tt)
        np.testing.assert_allclose(f1[i, :nn], r1, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(f2[i, :nn], r2, rtol=1e-10, atol=1e-12)
    cells.close()


def test_truncated_window_cells(cells_npz, orc):
    """t_start / t_end windows: shorter series, t(1) != 0, their own t_interp."""
    from transcriptioncycleinference_b200 import _lib, setup_cell
    from transcriptioncycleinference_b200.engine import Cells
    if _lib.device_count() < 1:
        pytest.skip("no CUDA device")
    co, cons = orc
    ts, m2s, p7s = [], [], []
    for c in range(0, 299, 15):
        o, n = int(cells_npz["off"][c]), int(cells_npz["N"][c])
        a, b, d = setup_cell.truncate(cells_npz["t"][o:o + n], cells_npz["ms2"][o:o + n], cells_npz["pp7"][o:o + n], 3.0, 22.5)
        ts.append(a); m2s.append(b); p7s.append(d)
    cells = Cells(ts, m2s, p7s)
    packed = dict(N=cells.N, off=cells.off, t=cells.t, ms2=cells.ms2, pp7=cells.pp7)
    rng = np.random.default_rng(1)
    n = 3000
    cid = rng.integers(0, cells.ncells, n).astype(np.int32)
    th = np.zeros((n, cells.ld))
    for i, c in enumerate(cid):
        th[i, :7 + int(cells.N[c])] = random_theta(rng, int(cells.N[c]), False)
        th[i, 2] = rng.uniform(0, 8)                   # onset inside / before the window
    ref = co.ss_batch(cons, packed, cid, th)
    for algo in (0, 1):
        got = cells.ss_batch(cid, th, algo=algo)
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-10
    for c in range(cells.ncells):
        assert np.array_equal(cells.t_interp(c), co.t_interp(ts[c]))
    cells.close()
