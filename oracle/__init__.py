"""CPU oracle for the transcription-cycle MCMC hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is product code.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import, link or execute it, and
only as the checker or the timed CPU baseline.  The product path
(transcriptioncycleinference_b200/) never imports this package and fails loudly when
its CUDA library is missing.

What it restates (reference file:line, paths relative to /root/reference):
  matlab_builtins.py  MATLAB colon / interp1 / nansum / mean / std(.,1) semantics
                      used at src/SumofSquaresFunction_TranscriptionCycleMCMC.m:29-30,55-56,64
  forward_literal.py  src/dependencies/ConstantElongationSim.m:1-69 (m x n matrix and all),
                      src/GetFluorFromPolPos.m:1-71,
                      src/SumofSquaresFunction_TranscriptionCycleMCMC.m:1-65
  setup.py            src/TranscriptionCycleMCMC.m:163-270 (per-cell constants) and
                      :276-312 (slicing, summaries, best-fit curves)
  dram.py             the DRAM loop of mcmcstat::mcmcrun as configured at
                      src/TranscriptionCycleMCMC.m:242-273
  tc_oracle.c         the same forward model, SS and DRAM loop in plain C (FP64),
                      OpenMP over cells; doubles as the timed CPU baseline

Parity pinning:
  * forward model: PINNED by the reference's own 299 golden vectors
    (TestScripts/28-Oct-2020-TestData.mat::MCMCplot.simMS2/simPP7 recomputed from
    ::MCMCresults.mean_*), tests/test_oracle_golden.py.
  * SS function (re-gridding, interp1, nansum): pinned STATISTICALLY by the
    probability-integral-transform test on the fixture's s2chain
    (tests/test_oracle_golden.py::test_s2chain_pit).
  * summaries: PINNED (299/299 MCMCresults recomputed from MCMCchain).
  * DRAM sampler: the algorithm lives in the third-party package mcmcstat
    (github.com/mjlaine/mcmcstat, NOT vendored and NOT version-pinned by the
    reference, README.md:5).  Beyond the items the 10-step fixture pins (qcov is a
    covariance, DR scale 5, sigma2 update law, chain(1,:)=x0) it is
    **parity unpinned**: restated from the published DRAM algorithm
    (Haario, Laine, Mira, Saksman 2006) and mcmcstat's documented defaults.
"""
