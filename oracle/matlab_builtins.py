"""MATLAB built-in semantics the reference's hot path relies on.  TEST INFRASTRUCTURE.

Each helper names the reference call site whose behaviour it restates
(paths relative to /root/reference).
"""
import math

import numpy as np

EPS = np.finfo(np.float64).eps


def colon(a, d, b):
    """MATLAB ``a:d:b`` for doubles (non-integer step allowed).

    Used at src/SumofSquaresFunction_TranscriptionCycleMCMC.m:30
    (``t_interp = t(1):dt:t(end)``).  Restates Moler's published colon algorithm:
    n = round((b-a)/d) with a tolerance test, right end point snapped to b when
    within 2*eps*max(|a|,|b|), and the vector filled symmetrically from both ends
    so that out(1)==a and out(end)==c exactly.
    """
    a = float(a); d = float(d); b = float(b)
    if d == 0 or (d > 0 and a > b) or (d < 0 and a < b) or any(map(math.isnan, (a, d, b))):
        return np.zeros(0)
    tol = 2.0 * EPS * max(abs(a), abs(b))
    sig = 1.0 if d > 0 else -1.0
    if a == math.floor(a) and d == 1:
        n = int(math.floor(b) - a)
    elif a == math.floor(a) and d == math.floor(d):
        q = math.floor(a / d)
        r = a - q * d
        n = int(math.floor((b - r) / d) - q)
    else:
        # MATLAB round(): half away from zero
        x = (b - a) / d
        n = int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))
        if sig * (a + n * d - b) > tol:
            n -= 1
    c = a + n * d
    if sig * (c - b) > -tol:
        c = b
    out = np.zeros(n + 1)
    k = np.arange(0, n // 2 + 1, dtype=np.float64)
    ki = k.astype(np.int64)
    out[ki] = a + k * d
    out[n - ki] = c - k * d
    if n % 2 == 0:
        out[n // 2] = (a + c) / 2
    return out


def mean_diff(t):
    """``mean(t(2:end)-t(1:(end-1)))`` — SumofSquares...m:29.  Sum of the N-1
    differences divided by N-1 (NOT (t(end)-t(1))/(N-1): last bits differ)."""
    t = np.asarray(t, dtype=np.float64)
    d = t[1:] - t[:-1]
    s = 0.0
    for x in d:          # sequential accumulation, as a scalar loop would
        s += float(x)
    return s / d.size


def interp1_linear(x, v, xq):
    """``interp1(x,v,xq)`` default method: linear, NaN outside [x(1),x(end)].

    SumofSquares...m:55-56.  x strictly increasing.  The last interval is closed
    on the right (xq == x(end) returns v(end)).
    """
    x = np.asarray(x, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    out = np.full(xq.shape, np.nan)
    n = x.size
    for q in range(xq.size):
        z = xq.flat[q]
        if not (z >= x[0] and z <= x[-1]):
            continue
        k = int(np.searchsorted(x, z, side="right")) - 1   # x[k] <= z < x[k+1]
        if k >= n - 1:
            k = n - 2
        s = (z - x[k]) / (x[k + 1] - x[k])
        out.flat[q] = v[k] + s * (v[k + 1] - v[k])
    return out


def nansum(x):
    """Statistics-Toolbox ``nansum``: NaNs contribute 0 — SumofSquares...m:64."""
    x = np.asarray(x, dtype=np.float64)
    return float(np.sum(x[~np.isnan(x)]))


def std_pop(x, axis=0):
    """``std(x,1)``: population normalisation (divide by n) —
    src/TranscriptionCycleMCMC.m:287-303."""
    return np.std(np.asarray(x, dtype=np.float64), axis=axis, ddof=0)


def find_first(mask):
    """``find(mask,1,'first')`` (0-based; None when empty)."""
    idx = np.flatnonzero(mask)
    return int(idx[0]) if idx.size else None


def find_last(mask):
    """``find(mask,1,'last')`` (0-based; None when empty)."""
    idx = np.flatnonzero(mask)
    return int(idx[-1]) if idx.size else None
