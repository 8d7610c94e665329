"""GPU parity of the batched ssfun and the forward curves, through the C ABI, against the oracle.
Tolerance: 1e-10 relative (north_star, FP64 path)."""
import numpy as np
import pytest

from conftest import golden_theta, random_theta

pytestmark = pytest.mark.gpu
TOL = 1e-10


def test_t_interp_matches_oracle(gpu_cells, orc):
    co, _ = orc
    for c in range(0, gpu_cells.ncells, 7):
        t, _, _ = gpu_cells.cell(c)
        assert np.array_equal(gpu_cells.t_interp(c), co.t_interp(t))


def test_forward_golden_vectors(gpu_cells, results_npz):
    """The reference's own 299 known-answer vectors (MCMCplot.simMS2/simPP7) on the raw grid."""
    g = results_npz
    th = gpu_cells.pad_theta([golden_theta(g, c) for c in range(299)])
    ms2, pp7 = gpu_cells.forward(np.arange(299), th, on_raw_grid=True)
    worst = 0.0
    for c in range(299):
        s = slice(int(g["off"][c]), int(g["off"][c + 1])); n = int(g["N"][c])
        worst = max(worst, np.max(np.abs(ms2[c, :n] - g["simMS2"][s]) / np.abs(g["simMS2"][s])),
                    np.max(np.abs(pp7[c, :n] - g["simPP7"][s]) / np.abs(g["simPP7"][s])))
    assert worst < 1e-12, worst


def test_ss_slow_elongation_long_ramps(gpu_cells, cells_npz, orc):
    """v -> 0: the ramp of the response spans tens to all of the lags (a regime real chains wander into: a cell whose
    posterior has a mode at v ~ 0.01).  The O(1) prefix-sum form of the ramp (counts K and first moments S) must agree
    with the literal m x n oracle there too."""
    co, cons = orc
    rng = np.random.default_rng(5)
    n = 12000
    cid = rng.integers(0, 299, n).astype(np.int32)
    th = np.zeros((n, gpu_cells.ld))
    for i, c in enumerate(cid):
        N = int(cells_npz["N"][c]); th[i, :7 + N] = random_theta(rng, N, False)
        th[i, 0] = 10.0 ** rng.uniform(-3.0, -0.3)           # v in [0.001, 0.5] kb/min
        th[i, 1] = rng.uniform(0.0, 20.0)                     # tau over its whole range: L = L0 + tau v
    ref = co.ss_batch(cons, cells_npz, cid, th)
    for algo in (1, 0):
        got = gpu_cells.ss_batch(cid, th, algo=algo)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        assert int((rel >= TOL).sum()) <= 1, (algo, rel.max())


@pytest.mark.parametrize("algo", [0, 1])
def test_ss_recorded_states(gpu_cells, chains_npz, cells_npz, orc, algo):
    """The 2 990 parameter vectors the reference itself recorded (MCMCchain) x their cell's data."""
    co, cons = orc
    th = np.nan_to_num(chains_npz["theta"].reshape(2990, -1))
    cid = np.repeat(np.arange(299, dtype=np.int32), 10)
    ref = co.ss_batch(cons, cells_npz, cid, th)
    got = gpu_cells.ss_batch(cid, th, algo=algo)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() < TOL, (rel.max(), np.argmax(rel))


@pytest.mark.parametrize("algo,wide,n", [(0, True, 20000), (1, True, 60000), (1, False, 60000), (0, False, 20000)])
def test_ss_random_theta(gpu_cells, cells_npz, orc, algo, wide, n):
    """theta uniform inside the sampler's bounds (seed 0) across all cells.  A polymerase within an
    ulp of the discontinuity at L may legitimately flip: such rows are counted and
    must be vanishingly rare; everything else is within 1e-10."""
    co, cons = orc
    rng = np.random.default_rng(0)
    cid = rng.integers(0, 299, n).astype(np.int32)
    th = np.zeros((n, gpu_cells.ld))
    for i, c in enumerate(cid):
        N = int(cells_npz["N"][c]); th[i, :7 + N] = random_theta(rng, N, wide)
    ref = co.ss_batch(cons, cells_npz, cid, th)
    got = gpu_cells.ss_batch(cid, th, algo=algo)
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
    flips = int((rel >= TOL).sum())
    assert flips <= 1, (flips, rel.max())
    assert np.median(rel) < 1e-13


def test_ss_algos_agree(gpu_cells, cells_npz):
    rng = np.random.default_rng(3)
    n = 50000
    cid = rng.integers(0, 299, n).astype(np.int32)
    th = np.zeros((n, gpu_cells.ld))
    for i, c in enumerate(cid):
        N = int(cells_npz["N"][c]); th[i, :7 + N] = random_theta(rng, N, i % 2 == 0)
    a = gpu_cells.ss_batch(cid, th, algo=0); b = gpu_cells.ss_batch(cid, th, algo=1)
    rel = np.abs(a - b) / np.abs(a)
    assert (rel >= TOL).sum() <= 1, rel.max()


def test_ss_edge_cases(gpu_cells, cells_npz, orc):
    """ton beyond the movie (no loading), v = 0, negative total rate, all within bounds."""
    co, cons = orc
    rng = np.random.default_rng(5)
    rows, cid = [], []
    for c in (0, 17, 298):
        N = int(cells_npz["N"][c])
        base = random_theta(rng, N, False)
        for mod in ("ton_late", "v_zero", "rate_neg", "tau_zero", "A_zero", "basal_high", "v_max"):
            th = base.copy()
            if mod == "ton_late": th[2] = 10.0
            if mod == "v_zero": th[0] = 0.0
            if mod == "rate_neg": th[6] = 0.0; th[7:] = -np.abs(th[7:])
            if mod == "tau_zero": th[1] = 0.0
            if mod == "A_zero": th[5] = 0.0
            if mod == "basal_high": th[3] = 50.0; th[4] = 50.0
            if mod == "v_max": th[0] = 10.0; th[1] = 20.0
            rows.append(th); cid.append(c)
    th = gpu_cells.pad_theta(rows); cid = np.array(cid, dtype=np.int32)
    ref = co.ss_batch(cons, cells_npz, cid, th)
    for algo in (0, 1):
        got = gpu_cells.ss_batch(cid, th, algo=algo)
        assert np.max(np.abs(got - ref) / np.abs(ref)) < TOL


def test_empty_batch_and_errors(gpu_cells):
    from transcriptioncycleinference_b200 import _lib
    assert gpu_cells.ss_batch(np.zeros(0, dtype=np.int32), np.zeros((0, gpu_cells.ld))).size == 0
    with pytest.raises(_lib.TcError):
        gpu_cells.ss_batch(np.array([299], dtype=np.int32), np.zeros((1, gpu_cells.ld)))
    with pytest.raises(_lib.TcError):
        gpu_cells.ss_batch(np.array([0], dtype=np.int32), np.zeros((1, 20)))


def test_ss_one_million_random_theta(gpu_cells, cells_npz, orc):
    """SURVEY.md 8(d)'s SS-parity volume: 10^6 theta drawn uniformly inside the bounds of TranscriptionCycleMCMC.m:242-254
    (seed 0) across all 299 cells, both algorithms, against the literal m x n oracle (all host cores), 1e-10 relative.
    A polymerase within an ulp of the discontinuity at L = L0 + tau v may flip: counted, reported, and
    bounded at 5 per million."""
    co, cons = orc
    rng = np.random.default_rng(0)
    n = 1_000_000
    cid = np.sort(rng.integers(0, 299, n)).astype(np.int32)           # proposals for a cell arrive together
    ld = gpu_cells.ld
    lo = np.concatenate([[0, 0, 0, 0, 0, 0, 0], -30 * np.ones(ld - 7)])
    hi = np.concatenate([[10, 20, 10, 50, 50, 1, 40], 30 * np.ones(ld - 7)])
    th = np.zeros((n, ld))
    for s in range(0, n, 100_000):                                   # (in slabs: 1.1 GB of theta in all)
        th[s:s + 100_000] = lo + (hi - lo) * rng.random((min(100_000, n - s), ld))
    npar = 7 + cells_npz["N"][cid]
    th[np.arange(ld)[None, :] >= npar[:, None]] = 0.0                # zero padding beyond each cell's npar
    ref = co.ss_batch(cons, cells_npz, cid, th)
    for algo in (1, 0):
        got = gpu_cells.ss_batch(cid, th, algo=algo)
        rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-300)
        flips = int((rel >= TOL).sum())
        print("algo %d: 10^6 evaluations, max rel %.3g, median %.3g, rows beyond 1e-10: %d" % (algo, rel.max(), np.median(rel), flips))
        assert flips <= 5, (algo, flips, rel.max())
        assert np.median(rel) < 1e-13
