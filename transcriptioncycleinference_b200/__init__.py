"""tcmcmc — B200-native (sm_100a) engine for the per-cell DRAM fit of the Liu et al. (2020)
transcription-cycle model: a drop-in for the hot path of GarciaLab/TranscriptionCycleInference
(TranscriptionCycleMCMC -> mcmcrun -> SumofSquaresFunction_TranscriptionCycleMCMC ->
ConstantElongationSim -> GetFluorFromPolPos).  All compute lives in libtcmcmc.so (CUDA, C ABI in
include/tcmcmc.h); this package is the host-side mirror of the reference interface."""
from . import _lib  # noqa: F401
from .constructs import DEFAULT_CONSTRUCT, get_construct, register_construct  # noqa: F401
from .engine import Cells  # noqa: F401

__all__ = ["Cells", "DEFAULT_CONSTRUCT", "get_construct", "register_construct"]
