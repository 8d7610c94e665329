#!/usr/bin/env python
"""bench.py — headline benchmark of the transcription-cycle MCMC hot path on B200.

Metric (BASELINE.json): MCMC chain-steps/s (and SS-likelihood evaluations/s) on the 299 cells of
TestData.mat.  One bench "step" = one complete DRAM fit of the workload (all chains, n_steps MCMC
steps each) by ONE launch of the device-resident sampler kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload config2|config3] [--n-steps S] [--n-burn B]

  config2 (default, BASELINE configs[1]): 299 cells x 1 chain, n_burn=10000, n_steps=200000.
  config3: 299 cells x 64 chains (19 136 chains), same steps.
With N > 1 (torchrun, one rank per GPU) every rank fits its own replica set of chains (weak
scaling: the path has no data-path collective; Philox streams are keyed by chain identity, never by
rank).  `value` = chain-steps of all ranks / max-over-ranks time.

--impl reference times the CPU restatement of the reference algorithm (oracle/, a C port: the
reference is MATLAB + un-vendored mcmcstat and cannot run here) on the box's host cores on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "cells.npz")
# dram__bytes_read.sum + dram__bytes_write.sum of ONE dram_kernel launch of the default workload (config 2, 200 000
# steps), from an ncu capture of this very command (profiles/traffic_dram_r1v.csv): 109 MB read + 8.0 GB written
# (distinct chain rows, and the part of the proposal factors / scatter matrices that L2 evicts; 3.8 GB at r1x: the
# write-back share moves with L2 residency, either way < 0.1 % of HBM bandwidth).  Other workloads: not captured.
MEASURED_TRAFFIC_BYTES = {("config2", 200000, 10000): 109115648 + 8000692480}
METRIC = "mcmc_chain_steps_per_s"
UNIT = "chain-steps/s"


def w_flop(N):
    """Algorithmic flops of one SS evaluation: 9 N(N-1)/2 + 30 N."""
    return 9.0 * N * (N - 1) / 2 + 30.0 * N


def w_op(N):
    """Algorithmic FP64-pipe instructions of one SS evaluation: 11 N(N-1)/2 + 20 N."""
    return 11.0 * N * (N - 1) / 2 + 20.0 * N


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"])
    ap.add_argument("--n-steps", type=int, default=200000)
    ap.add_argument("--n-burn", type=int, default=10000)
    ap.add_argument("--cpu-sample-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the secondary N = 400 leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(self.rows))


# ------------------------------------------------------------------------------- CPU baseline
def cpu_baseline(g, workload, n_burn, target_seconds, rank_offset=0):
    """Oracle (C port of the reference algorithm, literal m x n forward model + DRAM) on the host
    cores, OpenMP over chains (the parfor analogue).  Bounded sample: all 299 cells, one chain each,
    a reduced number of MCMC steps chosen to cost ~target_seconds of CPU time."""
    from oracle import c_oracle, forward_literal
    from transcriptioncycleinference_b200 import setup_cell

    cons = c_oracle.Construct.from_dict(forward_literal.CONSTRUCTS["P2P-MS2v5-LacZ-PP7v4"])
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ask the OS, not OpenMP)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1

    class HostCells:                      # the bits of engine.Cells that chain_inputs needs
        def __init__(s):
            s.ld = 7 + int(g["N"].max())

        def cell(s, c):
            o = int(g["off"][c]); n = int(g["N"][c])
            return g["t"][o:o + n], g["ms2"][o:o + n], g["pp7"][o:o + n]

    cc = np.arange(299, dtype=np.int32)
    inputs = setup_cell.chain_inputs(HostCells(), cc, np.random.default_rng(1))
    # calibrate: 40 steps
    opts = c_oracle.default_opts(40, 20)
    t0 = time.time()
    c_oracle.run_chains(cons, g, opts, 20, cc, *inputs, seed=1, nthreads=cores)
    per_step = (time.time() - t0) / (299 * 40)
    nsimu = int(max(100, min(20000, target_seconds / (per_step * 299))))
    nsimu = (nsimu // 100) * 100
    burn = nsimu // 2
    opts = c_oracle.default_opts(nsimu, burn)
    t0 = time.time()
    _, _, _, cnt = c_oracle.run_chains(cons, g, opts, burn, cc, *inputs, seed=2, nthreads=cores)
    dt = time.time() - t0
    return dict(value=299 * nsimu / dt, unit=UNIT, cores=cores, kind="port",
                sample="299 cells x 1 chain x %d MCMC steps (burnintime %d), literal m x n forward model, "
                       "C port of the reference algorithm (oracle/tc_oracle.c), OpenMP over chains" % (nsimu, burn),
                seconds=dt, ss_evals_per_s=float(cnt[:, 0].sum()) / dt)


def run_reference(args, g):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(g, args.workload, args.n_burn, max(2.0, args.cpu_sample_seconds / 3))
        if i >= args.warmup:
            vals.append(last)
    v = float(np.mean([x["value"] for x in vals]))
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean([x["seconds"] for x in vals]) * 1e3), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="TestData.mat cells (tests/golden/cells.npz), synthetic x0",
                impl="reference",
                config=dict(workload=args.workload + " (bounded sample)", cells=299, chains_per_cell=1),
                cpu_baseline=dict(value=v, unit=UNIT, cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                ss_evals_per_s=float(np.mean([x["ss_evals_per_s"] for x in vals])), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm
def run_ours(args, g):
    import torch
    import torch.distributed as dist
    from transcriptioncycleinference_b200 import _lib, setup_cell
    from transcriptioncycleinference_b200.engine import Cells

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libtcmcmc has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"            # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells = Cells.from_packed(g["N"], g["off"], g["t"], g["ms2"], g["pp7"], devices=(local,))
    per_cell = 1 if args.workload == "config2" else 64
    cc = np.repeat(np.arange(299, dtype=np.int32), per_cell)
    # chain identity: (cell, chain replica) — rank r owns replicas [r*per_cell, (r+1)*per_cell)
    rep = np.tile(np.arange(per_cell, dtype=np.uint64), 299) + np.uint64(rank * per_cell)
    uid = cc.astype(np.uint64) * np.uint64(1 << 20) + rep
    inputs = setup_cell.chain_inputs(cells, cc, np.random.default_rng(1000 + rank))
    opts = _lib.default_opts(nsimu=args.n_steps, burnintime=args.n_burn, n_burn=args.n_burn, store_chain=0,
                             seed=20201028, ngpus=1)
    opts.devices[0] = local
    h2d = sum(x.nbytes for x in inputs) + cc.nbytes + uid.nbytes
    ld = cells.ld
    d2h = cc.size * (2 * ld + 2) * 8 + cc.size * _lib.NCOUNTERS * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    peak_dfma, peak_clk = _lib.measure_fp64_peak(local)

    def one_fit():
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = cells.mcmc_run(opts, cc, *inputs, chain_uid=uid)          # host buffers in, summaries out
        return time.perf_counter() - t0, out

    for _ in range(args.warmup):
        one_fit()
    barrier()
    kernel_s, e2e_s, outs = [], [], []
    sampler = ClockSampler(local)
    with sampler:
        for _ in range(args.steps):
            dt, out = one_fit()
            e2e_s.append(dt); kernel_s.append(out["kernel_seconds"]); outs.append(out)
    barrier()
    tk, te = float(np.sum(kernel_s)), float(np.sum(e2e_s))
    if world > 1:
        tt = torch.tensor([tk, te], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tk, te = float(tt[0]), float(tt[1])
    nsteps_total = world * cc.size * args.n_steps * args.steps
    cnt = outs[-1]["counters"]
    Ns = g["N"][cc].astype(np.float64)
    npar = Ns + 7
    ss_evals = cnt[:, _lib.CNT_SS_EVALS].astype(np.float64)
    # algorithmic flops of one fit: SS evaluations + full-R proposal mat-vecs after the first
    # adaptation (2 proposals share one pass: 2 * npar^2/2 FMA) + covariance block updates
    # (npar^2/2 FMA per row) + Cholesky (npar^3/6 FMA) per adaptation     [SURVEY.md 8d]
    post = max(0, args.n_steps - args.n_burn)
    flops_ss = float(np.sum(ss_evals * w_flop(Ns)))
    flops_dram = float(np.sum(2.0 * post * npar ** 2 * 2 / 2 + args.n_steps * npar ** 2 + cnt[:, _lib.CNT_ADAPTATIONS] * npar ** 3 / 3))
    ops_ss = float(np.sum(ss_evals * w_op(Ns)))
    t_fit = tk / args.steps
    achieved = (flops_ss + flops_dram) / t_fit / 1e12
    peak_tf = 2 * peak_dfma / 1e12
    line = dict(
        metric=METRIC, value=nsteps_total / tk, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=tk / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
        data="TestData.mat cells (tests/golden/cells.npz: the reference's 299-cell dataset), synthetic x0",
        config=dict(workload="%s: 299 cells x %d chain(s)/cell per GPU, n_burn=%d, n_steps=%d, adaptint=100, DRAM, "
                             "construct P2P-MS2v5-LacZ-PP7v4" % (args.workload, per_cell, args.n_burn, args.n_steps),
                    chains_per_gpu=int(cc.size), l2="flushed between timed fits (256 MiB write)",
                    sharding="chains replicated per rank, no collective in the data path"),
        e2e=dict(value=nsteps_total / te, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
        gpu_launches=args.steps,
        ss_evals_per_s=float(world * ss_evals.sum() * args.steps / tk),
        ss_evals_per_chain_step=float(ss_evals.sum() / (cc.size * args.n_steps)),
        accept_rate=float((cnt[:, _lib.CNT_ACC_STAGE1] + cnt[:, _lib.CNT_ACC_STAGE2]).sum() / (cc.size * args.n_steps)),
        roofline=dict(bound="fp64_pipe", kernel="dram_kernel", achieved=achieved, peak=peak_tf, unit="TFLOP/s",
                      frac=achieved / peak_tf, traffic=MEASURED_TRAFFIC_BYTES.get((args.workload, args.n_steps, args.n_burn)),
                      traffic_unit="bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic_dram_r1v.csv); "
                                   "the kernel is bound by FP64-pipe latency, not HBM: this is 0.1 % of HBM bandwidth",
                      peak_source="measured DFMA micro-benchmark on this GPU (tc_measure_fp64_peak), SM clock %.0f MHz; "
                                  "MEASURED_PEAKS.json has no FP64 entry" % peak_clk,
                      algorithmic_flops_per_launch=flops_ss + flops_dram, flops_ss=flops_ss, flops_dram=flops_dram,
                      frac_ops_ss_only=ops_ss / t_fit / peak_dfma,
                      hbm_row_writeback_gbs=float(np.sum(npar) * args.n_steps * 8 / t_fit / 1e9)),
        clocks=sampler.summary(),
    )
    # secondary: the batched SS kernel alone (device-resident inputs), the compute-bound leg
    if rank == 0:
        line["ss_kernel"] = ss_kernel_leg(cells, g, peak_dfma, torch)
        if not args.no_config5:
            try:
                line["config5_scale"] = config5_leg(local)
            except Exception as e:                       # a secondary leg must not cost the headline line
                line["config5_scale"] = dict(error=repr(e))
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(g, args.workload, args.n_burn, args.cpu_sample_seconds)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config5_leg(device, ncells=1184, N=400, nsteps=2000):
    """Secondary: BASELINE config 5's series length (N = 400, npar = 407: the big layout of the sampler) at a single-GPU
    scale — synthetic cells from the forward model (transcriptioncycleinference_b200/synthetic.py), 8 chains per SM."""
    from transcriptioncycleinference_b200 import _lib, setup_cell, synthetic
    cells5, _ = synthetic.make_cells(ncells, N, devices=(device,))
    cc = np.arange(ncells, dtype=np.int32)
    inputs = setup_cell.chain_inputs(cells5, cc, np.random.default_rng(5))
    opts = _lib.default_opts(nsimu=nsteps, burnintime=nsteps // 2, n_burn=nsteps // 2, ngpus=1)
    opts.devices[0] = device
    ks = []
    for _ in range(3):
        ks.append(cells5.mcmc_run(opts, cc, *inputs)["kernel_seconds"])
    cells5.close()
    t = float(np.mean(ks[1:]))
    return dict(workload="%d synthetic cells x N=%d (npar=%d), 1 chain each, n_steps=%d, n_burn=%d" % (ncells, N, N + 7, nsteps, nsteps // 2),
                chain_steps_per_s=ncells * nsteps / t, ms=t * 1e3)


def ss_kernel_leg(cells, g, peak_dfma, torch):
    from transcriptioncycleinference_b200 import _lib
    rng = np.random.default_rng(0)
    per = 1024
    cid = np.repeat(np.arange(299, dtype=np.int32), per)
    th = np.zeros((cid.size, cells.ld))
    for c in range(299):
        N = int(g["N"][c]); m = slice(c * per, (c + 1) * per)
        lo = np.concatenate([[0.5, 0, 0, 0, 0, 0, 5], -8 * np.ones(N)])
        hi = np.concatenate([[4, 6, 6, 3, 3, 1, 25], 8 * np.ones(N)])
        th[m, :7 + N] = lo + (hi - lo) * rng.random((per, 7 + N))
    d_th = torch.from_numpy(th).cuda(); d_cid = torch.from_numpy(cid).cuda()
    d_out = torch.zeros(cid.size, dtype=torch.float64, device="cuda")
    res = {}
    for name, algo in (("toeplitz", _lib.ALGO_TOEPLITZ), ("pairs", _lib.ALGO_PAIRS)):
        ms = []
        for it in range(8):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            cells.ss_batch_device(cid.size, d_cid.data_ptr(), d_th.data_ptr(), cells.ld, d_out.data_ptr(), algo=algo)
            e1.record(); torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        t = float(np.mean(ms)) * 1e-3
        ops = float(np.sum(w_op(g["N"][cid].astype(np.float64))))
        res[name] = dict(evals_per_s=cid.size / t, ms=t * 1e3, batch=int(cid.size),
                         frac_of_fp64_peak_algorithmic_ops=ops / t / peak_dfma)
    res["note"] = ("device-resident inputs, %.0f MB of theta streamed from HBM per launch (> 126 MB L2); toeplitz = the O(N) "
                   "algorithm (ss_stream_kernel: operands staged in shared memory), pairs = the O(N^2) algorithm W_op counts "
                   "(ss_batch_kernel); frac = evals/s x W_op(N) / measured DFMA rate" % (th.nbytes / 1e6))
    return res


def main():
    args = parse()
    g = dict(np.load(GOLDEN))
    if args.impl == "reference":
        run_reference(args, g)
    else:
        run_ours(args, g)


if __name__ == "__main__":
    main()
