"""Synthetic scale-up workload (BASELINE config 5): cells with N time points generated from the
forward model with known parameters, for parameter-recovery checks and large-N throughput runs.

t = cumulative sum of dt ~ U(0.15, 0.35) min starting at 0 (irregular like the data); truth drawn from the x0 / prior ranges
of src/TranscriptionCycleMCMC.m:193-210 (v in [1,3], tau in [0,4], ton in [0,4], A in [0,1], MS2_basal, PP7_basal in [0,2],
R = 15, dR ~ N(0,3)); signals = the LIKELIHOOD'S OWN model — forward model on t_interp = t(1):dt:t(end), MS2 scaled by A,
interp1 back to the experimental times (SumofSquaresFunction_TranscriptionCycleMCMC.m:28-56, through tc_forward) — + N(0, sigma)
noise, so that the truth is the SS minimiser up to the noise (grid="raw" gives round 1's recipe: the plot call's model on the
raw grid, :307-309, which ssfun does not fit — rms residual at the truth 1.8 for sigma = 1); NaN mask Bernoulli(0.5) on MS2 and
(0.2) on PP7; numpy Philox generator, seed 20201028."""
import re

import numpy as np

from . import _lib
from .constructs import DEFAULT_CONSTRUCT
from .engine import Cells


def make_cells(ncells, N=400, seed=20201028, noise=1.0, construct=DEFAULT_CONSTRUCT, devices=(0,), batch=4096, grid="interp"):
    """-> (Cells resident on `devices`, truth [ncells, 7+N])"""
    rng = np.random.Generator(np.random.Philox(seed))
    t = np.concatenate([np.zeros((ncells, 1)), np.cumsum(rng.uniform(0.15, 0.35, (ncells, N - 1)), axis=1)], axis=1)
    truth = np.zeros((ncells, 7 + N))
    truth[:, 0] = rng.uniform(1, 3, ncells)
    truth[:, 1] = rng.uniform(0, 4, ncells)
    truth[:, 2] = rng.uniform(0, 4, ncells)
    truth[:, 3] = rng.uniform(0, 2, ncells)
    truth[:, 4] = rng.uniform(0, 2, ncells)
    truth[:, 5] = rng.uniform(0, 1, ncells)
    truth[:, 6] = 15.0
    truth[:, 7:] = rng.normal(0.0, 3.0, (ncells, N))
    ms2 = np.zeros((ncells, N)); pp7 = np.zeros((ncells, N))
    for b0 in range(0, ncells, batch):
        b1 = min(ncells, b0 + batch)
        z = list(np.zeros((b1 - b0, N)))
        for _ in range(64):
            # a grid whose t(1):dt:t(end) does not have N points is an error in the reference (R.*dt sizes, SumofSquares...m:30)
            # and in tc_cells_create (TC_EDIM, "... for cell K"): redraw that cell's time grid
            try:
                tmp = Cells(list(t[b0:b1]), z, z, construct=construct, devices=devices[:1])
                break
            except _lib.TcError as e:
                m = re.search(r"for cell (\d+)", str(e))
                if e.code != _lib.TC_EDIM or not m:
                    raise
                k = b0 + int(m.group(1))
                t[k, 1:] = np.cumsum(rng.uniform(0.15, 0.35, N - 1))
        m1, m2 = tmp.forward(np.arange(b1 - b0, dtype=np.int32), truth[b0:b1], on_raw_grid=(grid == "raw"))
        if grid == "raw":
            ms2[b0:b1] = m1[:, :N]; pp7[b0:b1] = m2[:, :N]
        else:
            for k in range(b1 - b0):                             # interp1(t_interp, model, t), :55-56
                tg = tmp.t_interp(k)
                ms2[b0 + k] = np.interp(t[b0 + k], tg, m1[k, :N]); pp7[b0 + k] = np.interp(t[b0 + k], tg, m2[k, :N])
        tmp.close()
    ms2 += rng.normal(0.0, noise, ms2.shape); pp7 += rng.normal(0.0, noise, pp7.shape)
    ms2[rng.random(ms2.shape) < 0.5] = np.nan
    pp7[rng.random(pp7.shape) < 0.2] = np.nan
    return Cells(list(t), list(ms2), list(pp7), construct=construct, devices=devices), truth


def recovery(truth, mean, std, idx=(0, 1, 2), nsig=3.0, keep=None):
    """fraction of cells whose true (v, tau, ton) lie within posterior mean +- nsig sigma, per parameter (keep: boolean mask
    of the cells that count, e.g. those whose chains agree)"""
    if keep is None:
        keep = np.ones(truth.shape[0], dtype=bool)
    if not keep.any():
        return [float("nan")] * len(idx)
    return [float(np.mean(np.abs(truth[keep, i] - mean[keep, i]) <= nsig * std[keep, i])) for i in idx]


def pool_chains(mean, std, nchains, n_rows):
    """Pool the `nchains` consecutive chains of every cell: pooled mean, pooled population std (within + between) and the
    Gelman-Rubin Rhat of (v, tau, ton) from the per-chain summaries.  -> (mean [ncells, ld], std, rhat [ncells, 3])"""
    from . import diagnostics
    ncells = mean.shape[0] // nchains
    mu = mean.reshape(ncells, nchains, -1); sd = std.reshape(ncells, nchains, -1)
    pm = mu.mean(axis=1)
    ps = np.sqrt((sd ** 2).mean(axis=1) + mu.var(axis=1))
    rh = np.stack([diagnostics.rhat_from_summaries(mu[c, :, :3], sd[c, :, :3], n_rows)[0] for c in range(ncells)])
    return pm, ps, rh
