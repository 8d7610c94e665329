"""Pin the CPU oracle against the reference's own known-answer material.
Everything here runs without a GPU."""
import numpy as np
import pytest
from scipy import stats

from conftest import DEFAULT_CONSTRUCT, golden_theta
from oracle import dram as pydram
from oracle import forward_literal as fl
from oracle import matlab_builtins as mb
from oracle import setup as osetup


def _cell(cells, c):
    o, n = int(cells["off"][c]), int(cells["N"][c])
    return cells["t"][o:o + n], cells["ms2"][o:o + n], cells["pp7"][o:o + n]


def test_c_oracle_forward_golden_all_299(orc, results_npz):
    """simMS2/simPP7 recomputed from the stored posterior means on the raw grid
    (TranscriptionCycleMCMC.m:307-309): 299 deterministic golden vectors."""
    co, cons = orc
    g = results_npz
    worst = 0.0
    for c in range(299):
        s = slice(int(g["off"][c]), int(g["off"][c + 1]))
        ms2, pp7 = co.model_on_grid(cons, golden_theta(g, c), g["t_plot"][s])
        worst = max(worst, np.max(np.abs(ms2 - g["simMS2"][s]) / np.abs(g["simMS2"][s])),
                    np.max(np.abs(pp7 - g["simPP7"][s]) / np.abs(g["simPP7"][s])))
    assert worst < 1e-12, worst


def test_python_literal_forward_golden_subset(results_npz):
    g = results_npz
    for c in range(0, 299, 23):
        s = slice(int(g["off"][c]), int(g["off"][c + 1]))
        ms2, pp7 = fl.model_on_grid(DEFAULT_CONSTRUCT, golden_theta(g, c), g["t_plot"][s])
        np.testing.assert_allclose(ms2, g["simMS2"][s], rtol=1e-12)
        np.testing.assert_allclose(pp7, g["simPP7"][s], rtol=1e-12)


def test_summaries_recomputed_from_chain(results_npz, chains_npz):
    """mean_*/sigma_* of MCMCresults = mean / std(.,1) of MCMCchain; mean_sigma = sqrt(mean(s2chain)),
    sigma_sigma = std(sqrt(s2chain),1)   (:286-303; the fixture run had n_burn = 1)."""
    g, ch = results_npz, chains_npz
    names = ["v", "tau", "ton", "MS2_basal", "PP7_basal", "A", "R"]
    for c in range(299):
        n = int(g["N"][c]); s = slice(int(g["off"][c]), int(g["off"][c + 1]))
        summ = osetup.summarise(ch["theta"][c][:, :7 + n], ch["s2chain"][c], 1)
        for k, nm in enumerate(names):
            assert abs(summ["mean"][k] - g["mean_" + nm][c]) <= 1e-12 * max(1, abs(g["mean_" + nm][c]))
            assert abs(summ["std"][k] - g["sigma_" + nm][c]) <= 1e-11 * max(1, abs(g["sigma_" + nm][c]))
        np.testing.assert_allclose(summ["mean"][7:], g["mean_dR"][s], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(summ["std"][7:], g["sigma_dR"][s], rtol=1e-10, atol=1e-13)
        assert abs(summ["mean_sigma"] - g["mean_sigma"][c]) < 1e-12 * g["mean_sigma"][c]
        assert abs(summ["sigma_sigma"] - g["sigma_sigma"][c]) < 1e-10 * max(1, g["sigma_sigma"][c])


def test_fixture_row1_is_x0(chains_npz):
    th = chains_npz["theta"]
    assert np.all(th[:, 0, 3] == 10) and np.all(th[:, 0, 4] == 5) and np.all(th[:, 0, 6] == 15)
    assert np.all(chains_npz["s2chain"][:, 0] == 1.0)


def test_s2chain_pit_pins_ss_function(orc, cells_npz, chains_npz):
    """sigma2(k) = (N0*S20 + SS(chain(k)))/chi2_nu with nu = N0 + 2*N_time: the probability-integral
    transform of the fixture's s2chain under the ORACLE's SS must be uniform.  This is what ties the
    oracle's complete SS function (t_interp, interp1, nansum) to the reference's output; nu from the
    non-NaN count is rejected."""
    co, cons = orc
    th = np.nan_to_num(chains_npz["theta"].reshape(2990, -1))
    cid = np.repeat(np.arange(299, dtype=np.int32), 10)
    ss = co.ss_batch(cons, cells_npz, cid, th).reshape(299, 10)
    s2 = chains_npz["s2chain"]
    N = cells_npz["N"].astype(np.float64)
    pit, pit_bad = [], []
    for c in range(299):
        nn = np.sum(~np.isnan(cells_npz["ms2"][cells_npz["off"][c]:cells_npz["off"][c + 1]])) + \
             np.sum(~np.isnan(cells_npz["pp7"][cells_npz["off"][c]:cells_npz["off"][c + 1]]))
        for k in range(1, 10):
            x = (1.0 + ss[c, k]) / s2[c, k]                       # ~ chi2(nu)
            pit.append(stats.chi2.cdf(x, 1 + 2 * N[c]))
            pit_bad.append(stats.chi2.cdf(x, 1 + nn))
    pit, pit_bad = np.array(pit), np.array(pit_bad)
    assert pit.size == 2691
    assert stats.kstest(pit, "uniform").pvalue > 0.01
    assert abs(pit.mean() - 0.5) < 0.02 and abs(pit.std() - 0.2887) < 0.02
    assert stats.kstest(pit_bad, "uniform").pvalue < 1e-6


def test_fixture_jump_variances_pin_qcov_and_drscale(chains_npz, cells_npz):
    """Accepted jumps in the dR block: variance 0.5 (J0 is a covariance) for stage 1 and 0.5/25 for
    the delayed-rejection stage (drscale = 5)."""
    th = chains_npz["theta"]
    v = []
    for c in range(299):
        n = int(cells_npz["N"][c])
        d = np.diff(th[c][:, 7:7 + n], axis=0)
        for row in d:
            if np.any(row != 0):
                v.append(row.var())
    v = np.array(v)
    big, small = v[v > 0.1], v[v <= 0.1]
    assert abs(big.mean() - 0.5) < 0.01
    assert abs(np.sqrt(0.5 / small.mean()) - 5.0) < 0.05


def test_colon_semantics():
    np.testing.assert_array_equal(mb.colon(0, 1, 5), np.arange(6.0))
    x = mb.colon(0, 0.1, 1)
    assert x.size == 11 and x[0] == 0 and x[-1] == 1.0
    assert mb.colon(1, 0.5, 0).size == 0
    x = mb.colon(0.3, 0.25, 2.31)       # end point not reached exactly
    assert x.size == 9 and abs(x[-1] - 2.3) < 1e-15
    # symmetric fill: second half counted down from the end point
    a, d, b = 0.0, 0.2521, 0.2521 * 119
    x = mb.colon(a, d, b)
    assert x.size == 120 and x[-1] == b and x[-2] == b - d and x[1] == a + d


def test_t_interp_all_cells(orc, cells_npz):
    co, _ = orc
    for c in range(299):
        t, _, _ = _cell(cells_npz, c)
        ti = fl.t_interp_of(t)
        assert ti.size == t.size and ti[0] == t[0] and ti[-1] == t[-1]
        assert np.array_equal(ti, co.t_interp(t))


def test_interp1_and_nansum():
    x = np.array([0.0, 1.0, 2.0]); v = np.array([0.0, 10.0, 30.0])
    out = mb.interp1_linear(x, v, np.array([-0.1, 0.0, 0.5, 1.0, 1.5, 2.0, 2.1]))
    assert np.isnan(out[0]) and np.isnan(out[-1])
    np.testing.assert_allclose(out[1:-1], [0, 5, 10, 20, 30])
    assert mb.nansum(np.array([1.0, np.nan, 2.0])) == 3.0
    assert mb.nansum(np.array([np.nan])) == 0.0


def test_python_ss_equals_c_ss(orc, cells_npz, chains_npz):
    co, cons = orc
    for c in (0, 101, 298):
        t, ms2, pp7 = _cell(cells_npz, c)
        for r in (0, 4, 9):
            th = chains_npz["theta"][c, r, :7 + t.size]
            a = fl.sum_of_squares(DEFAULT_CONSTRUCT, t, np.concatenate([ms2, pp7]), th)
            b = co.ss(cons, t, ms2, pp7, th)
            assert abs(a - b) <= 1e-12 * abs(a)


def test_multi_set_construct_per_set_clamp(orc):
    """Two loop sets: the basal clamp is applied after EACH set; python literal and C
    agree."""
    co, _ = orc
    d = dict(L_MS2=5.0, L_PP7=5.5, MS2_start=[0.1, 2.0], MS2_end=[1.0, 3.0], MS2_loopn=[24.0, 12.0],
             PP7_start=[3.2, 4.0], PP7_end=[3.9, 4.9], PP7_loopn=[24.0, 24.0])
    cons2 = co.Construct.from_dict(d)
    rng = np.random.default_rng(0)
    t = np.cumsum(np.concatenate([[0], rng.uniform(0.15, 0.35, 59)]))
    th = np.concatenate([[1.7, 2.0, 1.0, 4.0, 2.0, 0.4, 12.0], rng.normal(0, 3, 60)])
    a1, a2 = fl.model_on_grid(d, th, t)
    b1, b2 = co.model_on_grid(cons2, th, t)
    np.testing.assert_allclose(a1, b1, rtol=1e-13); np.testing.assert_allclose(a2, b2, rtol=1e-13)


def test_python_dram_equals_c_dram(orc, cells_npz):
    """Two independent restatements of the DRAM loop, same injected randomness: identical accept /
    reject flags through burn-in and covariance adaptation."""
    co, cons = orc
    t, ms2, pp7 = _cell(cells_npz, 7)
    rng = np.random.default_rng(1)
    x0 = osetup.initial_state(t, rng); J0 = osetup.proposal_variances(t)
    lo, hi, mu, sg = osetup.bounds_and_priors(t.size, x0)
    nsimu, burn = 400, 200
    st = pydram.make_streams(nsimu, x0.size, 1 + 2 * t.size, 7)
    # 200-row covariances are singular: whether chol(cov) "succeeds" is rounding noise, so both factor cov + qcovadj I
    rc = co.dram(cons, t, ms2, pp7, co.default_opts(nsimu, burn, qcovadj_always=1), x0, J0, lo, hi, mu, sg, streams=st)
    rp = pydram.dram(lambda th: co.ss(cons, t, ms2, pp7, th), x0, J0, lo, hi, mu, sg, 2 * t.size, nsimu, burn,
                     pydram.Recorded(**st), qcovadj_always=True)
    assert np.array_equal(rc["flags"], rp["flags"])
    np.testing.assert_allclose(rc["chain"], rp["chain"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(rc["s2chain"], rp["s2chain"], rtol=1e-9)
    assert rc["counters"][0] == rp["nss"] and rc["counters"][4] == 3


def test_c_dram_record_replay_and_invariants(orc, cells_npz):
    co, cons = orc
    t, ms2, pp7 = _cell(cells_npz, 33)
    rng = np.random.default_rng(2)
    x0 = osetup.initial_state(t, rng); J0 = osetup.proposal_variances(t)
    lo, hi, mu, sg = osetup.bounds_and_priors(t.size, x0)
    opts = co.default_opts(500, 200)
    r1 = co.dram(cons, t, ms2, pp7, opts, x0, J0, lo, hi, mu, sg, seed=5, record=True)
    st = {k: np.nan_to_num(v) for k, v in r1["streams"].items()}
    r2 = co.dram(cons, t, ms2, pp7, opts, x0, J0, lo, hi, mu, sg, streams=st)
    assert np.array_equal(r1["chain"], r2["chain"]) and np.array_equal(r1["s2chain"], r2["s2chain"])
    ch = r1["chain"]
    assert np.array_equal(ch[0], x0) and r1["s2chain"][0] == 1.0
    assert np.all(ch >= lo) and np.all(ch <= hi)
    # rejected steps repeat the previous row
    rej = (r1["flags"] & 1) == 0
    assert np.all(ch[1:][rej[1:]] == ch[:-1][rej[1:]])


def test_oracles_agree_at_config5_series_length(orc):
    """N = 400 (BASELINE config 5; npar = 407): the literal NumPy restatement and the C restatement of ssfun agree on a synthetic
    irregular grid with missing data, and the two DRAM restatements take the same decisions through a covariance adaptation
    (the GPU replay tests at N = 400 lean on the C one)."""
    co, cons = orc
    rng = np.random.default_rng(400)
    N = 400
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.15, 0.35, N - 1))])
    ms2 = rng.uniform(0, 3, N); pp7 = rng.uniform(0, 10, N)
    ms2[rng.random(N) < 0.5] = np.nan; pp7[rng.random(N) < 0.2] = np.nan
    assert co.t_interp(t).size == N
    for _ in range(3):
        th = np.concatenate([[rng.uniform(1, 3), rng.uniform(0, 4), rng.uniform(0, 4), rng.uniform(0, 2), rng.uniform(0, 2),
                              rng.uniform(0, 1), 15.0], rng.normal(0, 3, N)])
        a = fl.sum_of_squares(DEFAULT_CONSTRUCT, t, np.concatenate([ms2, pp7]), th)
        b = co.ss(cons, t, ms2, pp7, th)
        assert abs(a - b) <= 1e-12 * abs(a)
    x0 = osetup.initial_state(t, rng); J0 = osetup.proposal_variances(t)
    lo, hi, mu, sg = osetup.bounds_and_priors(N, x0)
    nsimu, burn = 150, 100
    st = pydram.make_streams(nsimu, x0.size, 1 + 2 * N, 11)
    rc = co.dram(cons, t, ms2, pp7, co.default_opts(nsimu, burn, qcovadj_always=1), x0, J0, lo, hi, mu, sg, streams=st)
    rp = pydram.dram(lambda th: co.ss(cons, t, ms2, pp7, th), x0, J0, lo, hi, mu, sg, 2 * N, nsimu, burn,
                     pydram.Recorded(**st), qcovadj_always=True)
    assert np.array_equal(rc["flags"], rp["flags"])
    np.testing.assert_allclose(rc["chain"], rp["chain"], rtol=0, atol=1e-7)
    assert rc["counters"][4] == 1                       # one adaptation with a 407 x 407 factorisation (step 100)


def test_dram_default_path_chol_first_then_blow(orc, cells_npz):
    """The defaults follow mcmcstat as SURVEY 3.2 / B.3 recall it: burn-in scaling on the CUMULATIVE rejection rate and
    chol(cov) first, chol(cov + qcovadj I) only when that fails.  (a) a chain that cannot move (low = upp = x0) has an
    exactly zero covariance: chol(cov) fails, the fallback gives R = sqrt(qcovadj) * adascale * I, counted as an
    adaptation, in both restatements; its burn-in scaling divides R by 10 at every decision (100 % rejections).
    (b) the two restatements still agree flag for flag on an ordinary chain with the fallback flag off / on."""
    co, cons = orc
    t, ms2, pp7 = _cell(cells_npz, 11)
    rng = np.random.default_rng(3)
    x0 = osetup.initial_state(t, rng); J0 = osetup.proposal_variances(t)
    lo, hi, mu, sg = osetup.bounds_and_priors(t.size, x0)
    o = co.default_opts(300, 200)
    assert o.burnin_cumulative == 1 and o.qcovadj_always == 0
    st = pydram.make_streams(300, x0.size, 1 + 2 * t.size, 9)
    rc = co.dram(cons, t, ms2, pp7, o, x0, J0, x0.copy(), x0.copy(), mu, sg, streams=st)
    rp = pydram.dram(lambda th: co.ss(cons, t, ms2, pp7, th), x0, J0, x0.copy(), x0.copy(), mu, sg, 2 * t.size, 300, 200,
                     pydram.Recorded(**st))
    assert np.array_equal(rc["flags"], rp["flags"]) and not np.any(rc["flags"] & 1)
    assert rc["counters"][4] == 2 and rc["counters"][5] == 0          # steps 200 and 300: fallback, not a failure
    assert np.all(rc["chain"] == x0)
