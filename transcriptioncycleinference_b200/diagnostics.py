"""Cross-chain convergence diagnostics from the per-chain summaries the sampler streams on the device
(mean, population std, number of kept rows) — no raw chains needed, so they are available for
BASELINE config 3 (299 cells x 64 chains, whose raw chains would be 3.9 TB).

Not part of the reference: `TranscriptionCycleMCMC.m` runs one chain per cell and
reports no diagnostic; with 'numChains' > 1 the extra variable `MCMCdiagnostics` is written next to
`MCMCresults` (the reference's layout is untouched).

Gelman-Rubin potential scale reduction (Gelman et al., BDA3 11.4, without chain splitting):
    W  = mean_c s_c^2                      within-chain variance (s_c^2 with n-1 normalisation)
    B  = n * var_c(mean_c)                 between-chain variance (m-1 normalisation)
    V+ = (n-1)/n * W + B/n
    Rhat = sqrt(V+ / W),   n_eff ~= m * n * V+ / B   (capped at m*n)
"""
import numpy as np


def rhat_from_summaries(means, stds, n):
    """means, stds: [m chains x p parameters] per-chain mean and POPULATION std (what tc_mcmc_run returns);
    n: kept rows per chain.  Returns (Rhat[p], n_eff[p]); NaN where a parameter did not move in any chain."""
    means = np.asarray(means, dtype=np.float64); stds = np.asarray(stds, dtype=np.float64)
    m = means.shape[0]
    if m < 2 or n < 2:
        raise ValueError("diagnostics need at least 2 chains of at least 2 rows")
    W = (stds ** 2).mean(axis=0) * n / (n - 1.0)
    B = n * means.var(axis=0, ddof=1)
    Vp = (n - 1.0) / n * W + B / n
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(Vp / W)
        neff = np.minimum(m * n * Vp / B, float(m * n))
    rhat[~(W > 0)] = np.nan
    neff[~(W > 0)] = np.nan
    return rhat, neff
