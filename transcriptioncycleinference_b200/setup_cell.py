"""Per-cell MCMC set-up — host-side mirror of src/TranscriptionCycleMCMC.m:163-255
(paths relative to the reference repo).  Product code: independent of oracle/."""
import numpy as np

# theta = [v, tau, ton, MS2_basal, PP7_basal, A, R, dR_1..dR_N]   (:210, :242-255)
PARAM_NAMES = ("v", "tau", "ton", "MS2_basal", "PP7_basal", "A", "R")
I_V, I_TAU, I_TON, I_MS2B, I_PP7B, I_A, I_R, I_DR = range(8)


def truncate(t, ms2, pp7, t_start=0.0, t_end=np.inf):
    """indStart = find(t >= t_start,1,'first'); indEnd = find(t < t_end,1,'last')   (:170-175)"""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    ms2 = np.asarray(ms2, dtype=np.float64).reshape(-1)
    pp7 = np.asarray(pp7, dtype=np.float64).reshape(-1)
    a = np.flatnonzero(t >= t_start)
    b = np.flatnonzero(t < t_end)
    if a.size == 0 or b.size == 0 or b[-1] < a[0]:
        return t[:0], ms2[:0], pp7[:0]
    return t[a[0]:b[-1] + 1], ms2[a[0]:b[-1] + 1], pp7[a[0]:b[-1] + 1]


def initial_state(N, rng, v0=None):
    """x0 (:193-210).  The reference draws from MATLAB's unseeded global stream; here `rng` is a
    numpy Generator so runs are reproducible.  v0 != None => loadPrevious (:193-198)."""
    v = (1.0 + 2.0 * rng.random()) if v0 is None else float(v0)
    ton0 = 4.0 * rng.random()
    A0 = rng.random()
    tau0 = 4.0 * rng.random()
    dR0 = rng.normal(0.0, 3.0, N)
    return np.concatenate([[v, tau0, ton0, 10.0, 5.0, A0, 15.0], dR0])


def proposal_variances(t, load_previous=False):
    """diag(J0) (:214-231); options.qcov = J0 (:266) so these are VARIANCES."""
    N = len(t)
    v_step = 0.0000001 if load_previous else 0.05
    return np.concatenate([[v_step, 0.1, t[-1] - t[-2], 1.0, 1.0, 0.05, 0.5], 0.5 * np.ones(N)])


def bounds_and_priors(N, x0, rate_prior_width=50.0, load_previous=False):
    """params cell array (:235-255) -> (low, upp, prior_mu, prior_sig)."""
    if load_previous:
        v_lo, v_hi = x0[0] - 0.00001, x0[0] + 0.00001
    else:
        v_lo, v_hi = 0.0, 10.0
    low = np.concatenate([[v_lo, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], -30.0 * np.ones(N)])
    upp = np.concatenate([[v_hi, 20.0, 10.0, 50.0, 50.0, 1.0, 40.0], 30.0 * np.ones(N)])
    mu = np.zeros(7 + N)
    sig = np.concatenate([np.full(7, np.inf), np.full(N, float(rate_prior_width))])
    return low, upp, mu, sig


def chain_inputs(cells, chain_cell, rng, rate_prior_width=50.0, v0=None):
    """Padded [nchains x ld] input arrays for Cells.mcmc_run.  v0: optional per-chain fixed
    elongation rate (loadPrevious)."""
    nch, ld = len(chain_cell), cells.ld
    th0 = np.zeros((nch, ld)); q = np.ones((nch, ld)); lo = np.zeros((nch, ld)); hi = np.zeros((nch, ld))
    mu = np.zeros((nch, ld)); sg = np.full((nch, ld), np.inf)
    for i, c in enumerate(chain_cell):
        t, _, _ = cells.cell(c)
        N = len(t)
        lp = v0 is not None
        x0 = initial_state(N, rng, None if not lp else v0[i])
        th0[i, :7 + N] = x0
        q[i, :7 + N] = proposal_variances(t, lp)
        a, b, m, s = bounds_and_priors(N, x0, rate_prior_width, lp)
        lo[i, :7 + N], hi[i, :7 + N], mu[i, :7 + N], sg[i, :7 + N] = a, b, m, s
    return th0, q, lo, hi, mu, sg
