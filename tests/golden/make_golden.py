#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's own .mat files.

Run HERE (the dev container, where /root/reference is mounted); the GPU box has
no /root/reference, so the tests, smoke() and bench.py only ever read the .npz
files this script writes next to itself.

Sources (all MAT v5, read with scipy.io.loadmat):
  /root/reference/TestScripts/TestData.mat                     -> cells.npz
  /root/reference/TestScripts/28-Oct-2020-TestData.mat         -> results.npz
  /root/reference/TestScripts/28-Oct-2020-TestData_RawChain.mat-> chains.npz

Layout of the packed ("ragged") arrays: cell c owns [off[c], off[c]+N[c]) of
every per-timepoint array, off = exclusive cumulative sum of N.

  cells.npz    N[299] int32, off[300] int64, t/ms2/pp7 [sum N] f64 (NaN = missing),
               name (str)                                  (README.md:11-16)
  results.npz  the 18 scalar MCMCresults fields as [299] f64 arrays (field order of
               TranscriptionCycleMCMC.m:151-155), mean_dR/sigma_dR packed [sum N],
               cell_index, ApprovedFits [299]; MCMCplot fields t_plot, MS2_plot,
               PP7_plot, simMS2, simPP7 packed [sum N]; DatasetName
  chains.npz   theta[299,10,7+Nmax] f64 in the sampler's own parameter order
               [v,tau,ton,MS2_basal,PP7_basal,A,R,dR_1..dR_N] (NaN padded beyond
               7+N[c]), s2chain[299,10]   (the shipped fixtures are a 10-step run,
               SURVEY.md section 0.1 #15)
"""
import os
import sys

import numpy as np
import scipy.io as sio

REF = "/root/reference/TestScripts"
HERE = os.path.dirname(os.path.abspath(__file__))

SCALAR_FIELDS = [
    "mean_v", "sigma_v", "mean_ton", "sigma_ton", "mean_A", "sigma_A", "mean_tau",
    "sigma_tau", "mean_MS2_basal", "sigma_MS2_basal", "mean_PP7_basal",
    "sigma_PP7_basal", "mean_R", "sigma_R", "mean_sigma", "sigma_sigma",
    "cell_index", "ApprovedFits",
]


def main():
    if not os.path.isdir(REF):
        sys.exit("reference fixtures not mounted at " + REF)

    d = sio.loadmat(os.path.join(REF, "TestData.mat"), mat_dtype=True)["data"]
    nc = d.shape[1]
    N = np.array([d[0, c]["time"].shape[1] for c in range(nc)], dtype=np.int32)
    off = np.zeros(nc + 1, dtype=np.int64)
    off[1:] = np.cumsum(N)
    t = np.concatenate([d[0, c]["time"][0] for c in range(nc)])
    ms2 = np.concatenate([d[0, c]["MS2"][0] for c in range(nc)])
    pp7 = np.concatenate([d[0, c]["PP7"][0] for c in range(nc)])
    name = str(d[0, 0]["name"][0])
    np.savez_compressed(os.path.join(HERE, "cells.npz"), N=N, off=off, t=t, ms2=ms2,
                        pp7=pp7, name=name)

    r = sio.loadmat(os.path.join(REF, "28-Oct-2020-TestData.mat"), mat_dtype=True)
    res, plot = r["MCMCresults"], r["MCMCplot"]
    out = {}
    for f in SCALAR_FIELDS:
        out[f] = np.array([float(res[0, c][f].squeeze()) for c in range(nc)])
    for f in ("mean_dR", "sigma_dR"):
        out[f] = np.concatenate([res[0, c][f].reshape(-1) for c in range(nc)])
    for f in ("t_plot", "MS2_plot", "PP7_plot", "simMS2", "simPP7"):
        out[f] = np.concatenate([plot[0, c][f].reshape(-1) for c in range(nc)])
    out["DatasetName"] = str(r["DatasetName"][0])
    out["field_order"] = np.array(res.dtype.names)
    out["plot_field_order"] = np.array(plot.dtype.names)
    np.savez_compressed(os.path.join(HERE, "results.npz"), N=N, off=off, **out)

    ch = sio.loadmat(os.path.join(REF, "28-Oct-2020-TestData_RawChain.mat"),
                     mat_dtype=True)["MCMCchain"]
    nrow = ch[0, 0]["v_chain"].shape[0]
    npmax = 7 + int(N.max())
    theta = np.full((nc, nrow, npmax), np.nan)
    s2 = np.zeros((nc, nrow))
    order = ["v_chain", "tau_chain", "ton_chain", "MS2_basal_chain", "PP7_basal_chain",
             "A_chain", "R_chain"]
    for c in range(nc):
        for k, f in enumerate(order):
            theta[c, :, k] = ch[0, c][f][:, 0]
        theta[c, :, 7:7 + N[c]] = ch[0, c]["dR_chain"]
        s2[c] = ch[0, c]["s2chain"][:, 0]
    np.savez_compressed(os.path.join(HERE, "chains.npz"), N=N, theta=theta, s2chain=s2,
                        chain_field_order=np.array(ch.dtype.names))
    for f in ("cells.npz", "results.npz", "chains.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
