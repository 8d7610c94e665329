"""Reporter-construct registry — the host-side mirror of the table in
src/GetFluorFromPolPos.m:18-45 (paths relative to the reference repo).

The reference's only user-extensible plug-in point is editing that if-block
(README.md:33-34).  Here a construct is a record with the same quantities under the same
names; `register_construct` is the equivalent of adding an `elseif` branch.  An unknown
name raises, like the reference (whose missing else-branch dies on an undefined variable).
"""
from . import _lib

_REGISTRY = {}


def register_construct(name, L_MS2, L_PP7, MS2_start, MS2_end, MS2_loopn, PP7_start, PP7_end,
                       PP7_loopn):
    """Lengths in kb.  L_* are the reporter lengths WITHOUT the dwell term; the engine adds
    tau*v (GetFluorFromPolPos.m:19-20).  start/end/loopn may be scalars or equal-length
    sequences (one entry per stem-loop set, :47)."""
    def vec(x):
        return [float(v) for v in (x if hasattr(x, "__len__") else [x])]
    rec = dict(L_MS2=float(L_MS2), L_PP7=float(L_PP7), MS2_start=vec(MS2_start), MS2_end=vec(MS2_end),
               MS2_loopn=vec(MS2_loopn), PP7_start=vec(PP7_start), PP7_end=vec(PP7_end),
               PP7_loopn=vec(PP7_loopn))
    n = len(rec["MS2_start"])
    if not (1 <= n <= _lib.MAX_SETS):
        raise ValueError("a construct needs 1..%d loop sets" % _lib.MAX_SETS)
    for k in ("MS2_end", "MS2_loopn", "PP7_start", "PP7_end", "PP7_loopn"):
        if len(rec[k]) != n:
            # the reference indexes the PP7 vectors with the MS2 loop index (:47,60-69)
            raise ValueError("construct %r: %s must have %d entries" % (name, k, n))
    _REGISTRY[name] = rec
    return rec


def get_construct(name):
    if isinstance(name, dict):
        return name
    if name not in _REGISTRY:
        raise NameError("Unrecognized construct %r: define it with register_construct() "
                        "(the reference fails on an undefined variable, GetFluorFromPolPos.m:47)" % (name,))
    return _REGISTRY[name]


def to_c(construct):
    rec = get_construct(construct)
    c = _lib.Construct()
    c.nsets = len(rec["MS2_start"])
    c.L_ms2, c.L_pp7 = rec["L_MS2"], rec["L_PP7"]
    for k in ("MS2_start", "MS2_end", "MS2_loopn", "PP7_start", "PP7_end", "PP7_loopn"):
        arr = getattr(c, k.lower())
        for i, x in enumerate(rec[k]):
            arr[i] = x
    return c


# The construct of Liu et al. (2020) — GetFluorFromPolPos.m:18-27
DEFAULT_CONSTRUCT = "P2P-MS2v5-LacZ-PP7v4"
register_construct(DEFAULT_CONSTRUCT, L_MS2=6.626, L_PP7=6.626, MS2_start=0.024, MS2_end=1.299,
                   MS2_loopn=24, PP7_start=4.292, PP7_end=5.758, PP7_loopn=24)
