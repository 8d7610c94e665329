"""Activates when tests/golden/mcmcstat_run.mat exists — the dump of ONE seeded mcmcrun call of real MATLAB + mcmcstat made by
baseline/run_reference.m (MATLAB is absent here, so the fixture is not shipped).  It pins against the
reference's own sampler what the reference's 10-step fixture cannot:
  * ssfun: the oracle's SS at every stored chain state equals mcmcstat's sschain (deterministic, 1e-10) — and the GPU's;
  * the sampler defaults the restatement assumes (N0, S20, drscale, adascale, qcovadj, burn-in scale, adaptint);
  * the adapted proposal: results.R'R = (cov(chain rows up to the last adaptation) [+ qcovadj I]) * adascale^2, which
    tells chol(cov)-first from always-regularised and pins the covariance recursion;
  * the sigma2 law (PIT of s2chain under (N0 S20 + ss)/chi2(N0 + 2 N))."""
import os

import numpy as np
import pytest

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcmcstat_run.mat")
pytestmark = pytest.mark.skipif(not os.path.exists(FIX), reason="no MATLAB + mcmcstat fixture (baseline/run_reference.m writes it)")


def _load():
    import scipy.io as sio
    m = sio.loadmat(FIX, mat_dtype=True, squeeze_me=True, struct_as_record=False)
    return m, m["results"]


def test_ssfun_equals_mcmcstat_sschain(orc, cells_npz):
    co, cons = orc
    m, _ = _load()
    c = int(m["cellNum"]) - 1
    o, N = int(cells_npz["off"][c]), int(cells_npz["N"][c])
    t, ms2, pp7 = (cells_npz[k][o:o + N] for k in ("t", "ms2", "pp7"))
    chain, ssc = np.atleast_2d(m["chain"]), np.asarray(m["sschain"]).reshape(-1)
    assert chain.shape[1] == 7 + N and np.array_equal(chain[0], np.asarray(m["x0"]).reshape(-1))
    rows = np.unique(np.linspace(0, chain.shape[0] - 1, 2000).astype(int))
    got = np.array([co.ss(cons, t, ms2, pp7, chain[r]) for r in rows])
    np.testing.assert_allclose(got, ssc[rows], rtol=1e-10)


def test_mcmcstat_defaults_match_the_restatement(orc):
    co, _ = orc
    _, r = _load()
    npar = int(r.npar)
    o = co.default_opts(int(r.nsimu), int(r.burnintime))
    assert int(r.adaptint) == o.adaptint == 100
    assert float(np.atleast_1d(r.drscale)[0]) == o.drscale == 5.0
    np.testing.assert_allclose(float(r.adascale), 2.4 / np.sqrt(npar), rtol=1e-12)
    assert float(r.qcovadj) == o.qcovadj
    assert float(np.atleast_1d(r.N0)[0]) == o.N0 and float(np.atleast_1d(r.S20)[0]) == o.S20
    if hasattr(r, "burnscale"):
        assert float(r.burnscale) == o.burnin_scale


def test_adapted_factor_is_chol_of_the_chain_covariance():
    m, r = _load()
    chain = np.atleast_2d(m["chain"])
    nsimu, adaptint = int(r.nsimu), int(r.adaptint)
    last = (nsimu // adaptint) * adaptint
    cov = np.cov(chain[:last].T)
    R = np.atleast_2d(r.R)
    got = R.T @ R / float(r.adascale) ** 2
    d = np.diag(got - cov)
    scale = np.sqrt(np.outer(np.diag(cov), np.diag(cov)))
    assert np.max(np.abs(got - cov - np.diag(d)) / scale) < 1e-6           # the covariance recursion
    # chol(cov) first => no regulariser on the diagonal; always-regularised => + qcovadj
    assert np.allclose(d, 0, atol=1e-10) or np.allclose(d, float(r.qcovadj), rtol=1e-3), "neither chol(cov) nor chol(cov + qcovadj I)"
    assert np.allclose(d, 0, atol=1e-10), "mcmcstat regularised a non-singular covariance: set qcovadj_always = 1 as the default"


def test_sigma2_law():
    from scipy import stats
    m, r = _load()
    s2, ssc = np.asarray(m["s2chain"]).reshape(-1), np.asarray(m["sschain"]).reshape(-1)
    N0, S20 = float(np.atleast_1d(r.N0)[0]), float(np.atleast_1d(r.S20)[0])
    nu = N0 + 2 * (np.atleast_2d(m["chain"]).shape[1] - 7)
    # mcmcstat stores the sigma2 drawn at the END of step k-1 in row k or the one of step k: accept whichever is uniform
    best = 0.0
    for a, b in ((s2[1:], ssc[1:]), (s2[1:], ssc[:-1])):
        u = stats.chi2.cdf((N0 * S20 + b) / a, nu)
        best = max(best, stats.kstest(u[::50], "uniform").pvalue)
    assert best > 0.01


@pytest.mark.gpu
def test_gpu_ssfun_equals_mcmcstat_sschain(gpu_cells, cells_npz):
    m, _ = _load()
    c = int(m["cellNum"]) - 1
    chain, ssc = np.atleast_2d(m["chain"]), np.asarray(m["sschain"]).reshape(-1)
    th = np.zeros((chain.shape[0], gpu_cells.ld)); th[:, :chain.shape[1]] = chain
    got = gpu_cells.ss_batch(np.full(chain.shape[0], c, dtype=np.int32), th)
    np.testing.assert_allclose(got, ssc, rtol=1e-10)
