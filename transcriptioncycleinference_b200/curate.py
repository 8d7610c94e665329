"""ApproveMCMCResults — headless, field-compatible curation of a results file.

The reference's src/ApproveMCMCResults.m is an interactive figure loop (approve / reject every single-cell fit by key press)
that reads fields the current TranscriptionCycleMCMC.m does not write (mean_dwell, dwell_chain, mean_R(2:end): SURVEY 0.1 #14),
from hard-coded S:\\ paths.  What the pipeline needs from it is its OUTPUT convention (:11-15, :335): MCMCresults(i).ApprovedFits
= 1 approved, 0 uncurated, -1 rejected, written back into the same .mat file, which 'loadPrevious' then carries into the next fit
(src/TranscriptionCycleMCMC.m:346-350).  This module applies that convention without a display, to the fields the current driver
DOES write:

    ApproveMCMCResults('file', path, 'approve', [1 5 9], 'reject', [2], 'maxRhat', 1.1, 'maxSigma', 3.0,
                       'minESS', 200, 'LoadPrevious', previous_results_file)

* 'approve' / 'reject': explicit 1-based positions in MCMCresults (what the key presses of the reference produce);
* automatic rules, applied to the fits that are still uncurated (0): 'maxRhat' — reject when MCMCdiagnostics.Rhat_max (several
  chains per cell, 'numChains') exceeds it, approve otherwise; 'minESS' — reject when the smallest ESS of the seven head
  parameters is below it (needs raw chains at fit time); 'maxSigma' — reject when mean_sigma (the fitted measurement noise)
  exceeds it;
* 'LoadPrevious' (the reference's option of the same name, :21-22): copy ApprovedFits from an earlier results file, matched
  on cell_index.
Every other variable of the file is written back unchanged.  Returns the vector of ApprovedFits."""
import numpy as np

OPTS = ("file", "approve", "reject", "maxrhat", "maxsigma", "miness", "loadprevious")


def ApproveMCMCResults(*varargin):
    import scipy.io as sio
    o = dict(file=None, approve=(), reject=(), maxrhat=None, maxsigma=None, miness=None, loadprevious=None)
    i = 0
    while i < len(varargin):
        k = varargin[i]
        if isinstance(k, str) and k.lower() in OPTS:
            if i + 1 >= len(varargin):
                raise IndexError("Index exceeds the number of array elements (option %r has no value)" % k)
            o[k.lower()] = varargin[i + 1]
            i += 2
        else:
            i += 1
    if o["file"] is None:
        raise ValueError("ApproveMCMCResults: 'file' (a results .mat written by TranscriptionCycleMCMC) is required")
    m = sio.loadmat(o["file"], mat_dtype=True)
    if "MCMCresults" not in m:
        raise KeyError("Reference to non-existent field 'MCMCresults'. (%s)" % o["file"])
    res = m["MCMCresults"]
    n = res.shape[1]
    app = np.array([float(np.asarray(res[0, k]["ApprovedFits"]).squeeze()) for k in range(n)])
    cell_index = np.array([int(np.asarray(res[0, k]["cell_index"]).squeeze()) for k in range(n)])
    if o["loadprevious"] is not None:
        p = sio.loadmat(o["loadprevious"], mat_dtype=True)["MCMCresults"]
        prev = {int(np.asarray(p[0, k]["cell_index"]).squeeze()): float(np.asarray(p[0, k]["ApprovedFits"]).squeeze())
                for k in range(p.shape[1])}
        for k in range(n):
            if cell_index[k] in prev:
                app[k] = prev[cell_index[k]]
    for k in np.atleast_1d(np.asarray(o["approve"], dtype=int)):
        app[k - 1] = 1.0
    for k in np.atleast_1d(np.asarray(o["reject"], dtype=int)):
        app[k - 1] = -1.0
    auto = app == 0.0
    verdict = np.zeros(n)
    if o["maxrhat"] is not None or o["miness"] is not None:
        if "MCMCdiagnostics" not in m:
            raise KeyError("'maxRhat' / 'minESS' need MCMCdiagnostics: fit with 'numChains' > 1")
        d = m["MCMCdiagnostics"]
        by_cell = {int(np.asarray(d[0, k]["cell_index"]).squeeze()): d[0, k] for k in range(d.shape[1])}
        for k in range(n):
            dg = by_cell.get(cell_index[k])
            if dg is None:
                continue
            ok = True
            if o["maxrhat"] is not None:
                ok &= float(np.asarray(dg["Rhat_max"]).squeeze()) <= float(o["maxrhat"])
            if o["miness"] is not None and np.asarray(dg["ESS"]).size:
                ok &= float(np.nanmin(np.asarray(dg["ESS"]).reshape(-1)[:7])) >= float(o["miness"])
            verdict[k] = 1.0 if ok else -1.0
    if o["maxsigma"] is not None:
        for k in range(n):
            if float(np.asarray(res[0, k]["mean_sigma"]).squeeze()) > float(o["maxsigma"]):
                verdict[k] = -1.0
            elif verdict[k] == 0.0:
                verdict[k] = 1.0
    app[auto] = np.where(verdict[auto] != 0.0, verdict[auto], app[auto])
    for k in range(n):
        res[0, k]["ApprovedFits"] = np.float64(app[k])
    out = {k: v for k, v in m.items() if not k.startswith("__")}
    out["MCMCresults"] = res
    sio.savemat(o["file"], out)                              # m.MCMCresults = MCMCresults  (:335)
    return app
