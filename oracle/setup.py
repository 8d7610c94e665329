"""Per-cell MCMC set-up and post-processing of the reference driver, restated.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows src/TranscriptionCycleMCMC.m
(paths relative to /root/reference)."""
import numpy as np

from .matlab_builtins import find_first, find_last, std_pop
from . import forward_literal as fl


def truncate(t, ms2, pp7, t_start=0.0, t_end=np.inf):
    """:170-175  indStart = find(t>=t_start,1,'first'); indEnd = find(t<t_end,1,'last')."""
    t = np.asarray(t, dtype=np.float64)
    i0, i1 = find_first(t >= t_start), find_last(t < t_end)
    if i0 is None or i1 is None:
        return t[:0], np.asarray(ms2)[:0], np.asarray(pp7)[:0]
    return t[i0:i1 + 1], np.asarray(ms2)[i0:i1 + 1], np.asarray(pp7)[i0:i1 + 1]


def initial_state(t, rng, v0=None):
    """:193-210.  rng supplies rand()/normrnd; v0 given => loadPrevious."""
    N = len(t)
    v = (1 + 2 * rng.random()) if v0 is None else v0
    ton0 = 4 * rng.random()
    A0 = rng.random()
    tau0 = 4 * rng.random()
    dR0 = rng.normal(0.0, 3.0, N)
    return np.concatenate([[v, tau0, ton0, 10.0, 5.0, A0, 15.0], dR0])


def proposal_variances(t, load_previous=False):
    """J0 diagonal :217-231 (these are VARIANCES: options.qcov = J0, :266)."""
    N = len(t)
    v_step = 1e-7 if load_previous else 0.05
    return np.concatenate([[v_step, 0.1, t[-1] - t[-2], 1.0, 1.0, 0.05, 0.5], 0.5 * np.ones(N)])


def bounds_and_priors(N, x0, rate_prior_width=50.0, load_previous=False):
    """params cell array :235-255 -> (low, upp, prior_mu, prior_sig)."""
    if load_previous:
        v_lo, v_hi = x0[0] - 0.00001, x0[0] + 0.00001
    else:
        v_lo, v_hi = 0.0, 10.0
    low = np.concatenate([[v_lo, 0, 0, 0, 0, 0, 0], -30.0 * np.ones(N)])
    upp = np.concatenate([[v_hi, 20, 10, 50, 50, 1, 40], 30.0 * np.ones(N)])
    mu = np.zeros(7 + N)
    sig = np.concatenate([np.full(7, np.inf), np.full(N, float(rate_prior_width))])
    return low, upp, mu, sig


def summarise(chain, s2chain, n_burn):
    """:276-303.  chain(n_burn:end,:) -> means and population stds; s2chain untrimmed."""
    c = chain[n_burn - 1:]
    mean, sd = c.mean(axis=0), std_pop(c, axis=0)
    return dict(mean=mean, std=sd, mean_sigma=float(np.sqrt(np.mean(s2chain))),
                sigma_sigma=float(std_pop(np.sqrt(s2chain))))


def best_fit_curves(construct, mean_theta, t_raw):
    """:307-309 — forward model at the posterior means on the RAW grid, no interp1."""
    return fl.model_on_grid(construct, mean_theta, t_raw)
