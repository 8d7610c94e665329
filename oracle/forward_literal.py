"""Literal NumPy restatement of the reference's forward model and SS function.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows the .m files line by line, m x n polymerase-position matrix and all — no
cohort shortcut, no closed-form positions.  Paths relative to /root/reference.
"""
import numpy as np

from .matlab_builtins import colon, interp1_linear, mean_diff, nansum

# Construct table: src/GetFluorFromPolPos.m:18-30.  Lengths are the base lengths to
# which tau*v is added (:19-20); start/end/loopn are vectors, one entry per loop set.
CONSTRUCTS = {
    "P2P-MS2v5-LacZ-PP7v4": dict(
        L_MS2=6.626, L_PP7=6.626,
        MS2_start=[0.024], MS2_end=[1.299], MS2_loopn=[24.0],
        PP7_start=[4.292], PP7_end=[5.758], PP7_loopn=[24.0],
    ),
}


def constant_elongation_sim(v, ton, R, t):
    """src/dependencies/ConstantElongationSim.m:1-69.  Returns the m x n matrix."""
    R = np.array(R, dtype=np.float64)[:-1]            # :33  drop last rate
    R[R < 0] = 0.0                                     # :36
    t = np.asarray(t, dtype=np.float64)
    m = t.size                                         # :39
    dt = np.full(m - 1, np.nan)
    for i in range(m - 1):                             # :43-45
        dt[i] = t[i + 1] - t[i]
    n = int(np.floor(np.sum(R * dt)))                  # :47
    n = max(n, 0)
    x = np.zeros((m, n))                               # :50
    counter = 0.0                                      # :53
    for i in range(m - 1):                             # :56
        if t[i] < ton:                                 # :57-58
            continue
        counter = counter + R[i] * dt[i]               # :60
        k = int(np.floor(counter))                     # :61  k = 1:floor(counter)
        if k > n:      # MATLAB would grow the matrix; sum() vs running-sum rounding only
            x = np.concatenate([x, np.zeros((m, k - n))], axis=1)
            n = k
        x[i + 1, :k] = x[i, :k] + v * dt[i]            # :64
        # :65 `x(x(i+1,k)<0)=0` is a no-op for v >= 0; not replicated
    return x


def get_fluor_from_pol_pos(construct, PolPos, v, tau, MS2_basal, PP7_basal):
    """src/GetFluorFromPolPos.m:1-71 (strict comparisons, per-set basal clamp)."""
    if isinstance(construct, str):
        if construct not in CONSTRUCTS:
            # the .m has no else-branch: an unknown name dies on an undefined variable
            raise NameError("Unrecognized function or variable 'MS2_start'.")
        construct = CONSTRUCTS[construct]
    c = construct
    L_MS2 = c["L_MS2"] + tau * v                       # :19
    L_PP7 = c["L_PP7"] + tau * v                       # :20
    m = PolPos.shape[0]
    MS2 = np.zeros(m)                                  # :29 (scalar 0 broadcast)
    PP7 = np.zeros(m)
    for i in range(len(c["MS2_start"])):               # :47
        s, e, ln = c["MS2_start"][i], c["MS2_end"][i], c["MS2_loopn"][i]
        fv = ln / 24                                   # :48
        mp = np.zeros(PolPos.shape)                    # :49
        mp[(PolPos > e) & (PolPos < L_MS2)] = fv       # :50
        fr = (PolPos > s) & (PolPos < e)               # :51
        mp[fr] = (PolPos[fr] - s) * fv / (e - s)       # :52
        MS2 = MS2 + np.sum(mp, axis=1)                 # :54
        MS2[MS2 < MS2_basal] = MS2_basal               # :57 (inside the loop)
        s, e, ln = c["PP7_start"][i], c["PP7_end"][i], c["PP7_loopn"][i]
        fv = ln / 24                                   # :60
        mp = np.zeros(PolPos.shape)
        mp[(PolPos > e) & (PolPos < L_PP7)] = fv       # :62
        fr = (PolPos > s) & (PolPos < e)               # :63
        mp[fr] = (PolPos[fr] - s) * fv / (e - s)       # :64
        PP7 = PP7 + np.sum(mp, axis=1)                 # :66
        PP7[PP7 < PP7_basal] = PP7_basal               # :69
    return MS2, PP7


def model_on_grid(construct, theta, tgrid):
    """[A*MS2, PP7] on an arbitrary grid: ConstantElongationSim + GetFluorFromPolPos
    + MS2 scaling — the body shared by SumofSquares...m:49-51 (tgrid = t_interp) and
    TranscriptionCycleMCMC.m:307-309 (tgrid = raw data.xdata)."""
    theta = np.asarray(theta, dtype=np.float64)
    v, tau, ton, b_ms2, b_pp7, A, R = theta[:7]
    dR = theta[7:]
    x = constant_elongation_sim(v, ton, R + dR, tgrid)
    MS2, PP7 = get_fluor_from_pol_pos(construct, x, v, tau, b_ms2, b_pp7)
    return A * MS2, PP7


def t_interp_of(t):
    """SumofSquares...m:29-30."""
    t = np.asarray(t, dtype=np.float64)
    dt = mean_diff(t)
    return colon(t[0], dt, t[-1])


def sum_of_squares(construct, t, ydata, theta):
    """src/SumofSquaresFunction_TranscriptionCycleMCMC.m:1-65.
    t = data.xdata (1 x N), ydata = [MS2, PP7] (1 x 2N, NaN = missing)."""
    t = np.asarray(t, dtype=np.float64)
    ti = t_interp_of(t)                                # :28-30
    MS2, PP7 = model_on_grid(construct, theta, ti)     # :49-51
    MS2 = interp1_linear(ti, MS2, t)                   # :55
    PP7 = interp1_linear(ti, PP7, t)                   # :56
    sim = np.concatenate([MS2, PP7])                   # :57
    res = np.asarray(ydata, dtype=np.float64) - sim    # :61
    return nansum(res ** 2)                            # :64
