"""ctypes binding of oracle/_build/libtcoracle.so (the C restatement).
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libtcoracle.so")
MAX_SETS = 8

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class Construct(C.Structure):
    _fields_ = [("nsets", C.c_int), ("L_ms2", C.c_double), ("L_pp7", C.c_double)] + [
        (n, C.c_double * MAX_SETS)
        for n in ("ms2_start", "ms2_end", "ms2_loopn", "pp7_start", "pp7_end", "pp7_loopn")
    ]

    @classmethod
    def from_dict(cls, d):
        c = cls()
        c.nsets = len(d["MS2_start"])
        c.L_ms2, c.L_pp7 = d["L_MS2"], d["L_PP7"]
        for k in ("MS2_start", "MS2_end", "MS2_loopn", "PP7_start", "PP7_end", "PP7_loopn"):
            arr = getattr(c, k.lower())
            for i, x in enumerate(d[k]):
                arr[i] = float(x)
        return c


class DramOpts(C.Structure):
    _fields_ = [
        ("nsimu", C.c_int), ("burnintime", C.c_int), ("adaptint", C.c_int), ("ntry", C.c_int),
        ("updatesigma", C.c_int), ("burnin_cumulative", C.c_int), ("qcovadj_always", C.c_int),
        ("drscale", C.c_double), ("adascale", C.c_double), ("qcovadj", C.c_double),
        ("burnin_scale", C.c_double), ("N0", C.c_double), ("S20", C.c_double),
        ("sigma2_0", C.c_double), ("Nobs", C.c_double),
    ]


def default_opts(nsimu, burnintime, Nobs=0.0, **kw):
    o = DramOpts(nsimu=nsimu, burnintime=burnintime, adaptint=100, ntry=2, updatesigma=1,
                 burnin_cumulative=1, qcovadj_always=0, drscale=5.0, adascale=0.0, qcovadj=1e-8, burnin_scale=10.0,
                 N0=1.0, S20=1.0, sigma2_0=1.0, Nobs=Nobs)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build(force=False):
    src = os.path.join(HERE, "tc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(LIB_PATH)):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_t_interp.argtypes = [C.c_int, _dp, _dp]
        L.orc_t_interp.restype = C.c_int
        L.orc_model_on_grid.argtypes = [C.POINTER(Construct), C.c_int, _dp, _dp, _dp, _dp]
        L.orc_model_on_grid.restype = None
        L.orc_ss.argtypes = [C.POINTER(Construct), C.c_int, _dp, _dp, _dp, _dp]
        L.orc_ss.restype = C.c_double
        L.orc_ss_batch.argtypes = [C.POINTER(Construct), _ip, _lp, _dp, _dp, _dp, C.c_longlong, _ip,
                                   _dp, C.c_int, _dp, C.c_int]
        L.orc_ss_batch.restype = None
        vp = C.c_void_p
        L.orc_dram.argtypes = [C.POINTER(Construct), C.c_int, _dp, _dp, _dp, C.POINTER(DramOpts),
                               _dp, _dp, _dp, _dp, _dp, _dp, vp, vp, vp, vp, vp, C.c_uint64,
                               vp, vp, vp, vp, vp, _dp, _dp, vp, vp, _lp]
        L.orc_dram.restype = C.c_int
        L.orc_run_chains.argtypes = [C.POINTER(Construct), _ip, _lp, _dp, _dp, _dp,
                                     C.POINTER(DramOpts), C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp,
                                     _dp, _dp, C.c_int, C.c_uint64, _dp, _dp, _dp, _lp, C.c_int]
        L.orc_run_chains.restype = None
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def t_interp(t):
    t = _c(t)
    out = np.zeros_like(t)
    n = lib().orc_t_interp(t.size, t, out)
    if n != t.size:
        raise ValueError("numel(t_interp) != N (%d vs %d)" % (n, t.size))
    return out


def model_on_grid(construct, theta, tgrid):
    tgrid = _c(tgrid)
    theta = _c(theta)
    assert theta.size == 7 + tgrid.size
    a, b = np.zeros_like(tgrid), np.zeros_like(tgrid)
    lib().orc_model_on_grid(C.byref(construct), tgrid.size, tgrid, theta, a, b)
    return a, b


def ss(construct, t, ms2, pp7, theta):
    t = _c(t)
    return lib().orc_ss(C.byref(construct), t.size, t, _c(ms2), _c(pp7), _c(theta))


def ss_batch(construct, cells, cell_id, theta, nthreads=0):
    """cells: dict with N (int32), off (int64), t, ms2, pp7; theta [nbatch, ld]."""
    theta = _c(theta)
    cell_id = np.ascontiguousarray(cell_id, dtype=np.int32)
    out = np.zeros(theta.shape[0])
    lib().orc_ss_batch(C.byref(construct), np.ascontiguousarray(cells["N"], dtype=np.int32),
                       np.ascontiguousarray(cells["off"], dtype=np.int64), _c(cells["t"]),
                       _c(cells["ms2"]), _c(cells["pp7"]), theta.shape[0], cell_id, theta,
                       theta.shape[1], out, nthreads)
    return out


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dram(construct, t, ms2, pp7, opts, theta0, qcov_diag, low, upp, pmu, psig, streams=None,
         seed=0, record=False):
    """One chain.  streams = dict(z1,u1,z2,u2,chi2) to inject randomness, else internal RNG.
    Returns dict(chain, s2chain, sschain, flags, counters[, rec streams])."""
    t = _c(t)
    N = t.size
    npar = 7 + N
    ns = opts.nsimu
    o = DramOpts.from_buffer_copy(opts)
    if o.Nobs == 0:
        o.Nobs = 2.0 * N
    chain = np.zeros((ns, npar)); s2 = np.zeros(ns); ssc = np.zeros(ns)
    flags = np.zeros(ns, dtype=np.int32); cnt = np.zeros(6, dtype=np.int64)
    ins = [None] * 5
    if streams is not None:
        ins = [_c(streams[k]) for k in ("z1", "u1", "z2", "u2", "chi2")]
        assert ins[0].shape == (ns, npar) and ins[2].shape == (ns, npar)
    rec = [None] * 5
    if record:
        rec = [np.full((ns, npar), np.nan), np.full(ns, np.nan), np.full((ns, npar), np.nan),
               np.full(ns, np.nan), np.full(ns, np.nan)]
    rc = lib().orc_dram(C.byref(construct), N, t, _c(ms2), _c(pp7), C.byref(o), _c(theta0),
                        _c(qcov_diag), _c(low), _c(upp), _c(pmu), _c(psig),
                        *[_vp(a) for a in ins], seed, *[_vp(a) for a in rec],
                        chain, s2, _vp(ssc), _vp(flags), cnt)
    if rc != 0:
        raise RuntimeError("orc_dram: ss(x0) not finite")
    out = dict(chain=chain, s2chain=s2, sschain=ssc, flags=flags, counters=cnt)
    if record:
        out["streams"] = dict(zip(("z1", "u1", "z2", "u2", "chi2"), rec))
    return out


def run_chains(construct, cells, opts, n_burn, chain_cell, theta0, qcov_diag, low, upp, pmu, psig,
               seed=0, nthreads=0):
    """OpenMP over chains; returns (mean, std, sig[nch,2], counters[nch,6])."""
    theta0 = _c(theta0)
    nch, ld = theta0.shape
    chain_cell = np.ascontiguousarray(chain_cell, dtype=np.int32)
    mean = np.zeros((nch, ld)); std = np.zeros((nch, ld)); sig = np.zeros((nch, 2))
    cnt = np.zeros((nch, 6), dtype=np.int64)
    lib().orc_run_chains(C.byref(construct), np.ascontiguousarray(cells["N"], dtype=np.int32),
                         np.ascontiguousarray(cells["off"], dtype=np.int64), _c(cells["t"]),
                         _c(cells["ms2"]), _c(cells["pp7"]), C.byref(opts), n_burn, nch, chain_cell,
                         theta0, _c(qcov_diag), _c(low), _c(upp), _c(pmu), _c(psig), ld, seed,
                         mean, std, sig, cnt, nthreads)
    return mean, std, sig, cnt


def max_threads():
    return lib().orc_max_threads()
