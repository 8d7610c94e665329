#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches csv into the committed text summaries under profiles/."""
import csv, io, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)
KEYS = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_op_dmma", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "launch__shared_mem_per_block", "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct", "lts__t_bytes.sum "]
def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2]
for name in ("dram", "ss", "c5", "warp"):
    rep = "gpurun_out/prof_%s_%s.ncu-rep" % (name, R)
    if not os.path.exists(rep):
        continue
    h, u, v = raw(rep)
    with open("profiles/ncu_%s_%s.txt" % (name, R), "w") as f:
        f.write("# ncu --set full --clock-control none, kernel %s, report %s (one launch)\n" % (name, os.path.basename(rep)))
        for a, b, c in zip(h, u, v):
            if a in ("Kernel Name", "Block Size", "Grid Size") or any(k.strip() in a for k in KEYS):
                f.write("%-90s %-14s %s\n" % (a, b, c))
        f.write("\n# by source function (warp instructions executed, stall samples)\n")
        f.write(subprocess.run([sys.executable, "scripts/ncu_wk_funcs.py" if name == "warp" else "scripts/ncu_funcs.py", rep], capture_output=True, text=True).stdout)
        f.write("\n# top source lines by stall samples\n")
        f.write(subprocess.run([sys.executable, "scripts/ncu_lines.py", rep, "25"], capture_output=True, text=True).stdout)
lst = "gpurun_out/launches_%s.csv" % R
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if r and r[0].isdigit()]
    hdr = None
    for r in csv.reader(open(lst)):
        if r and r[0] == "ID": hdr = r
    agg = {}
    for r in rows:
        d = dict(zip(hdr, r))
        k = d["Kernel Name"].split("(")[0]
        val = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += val * scale
    tot = sum(v[1] for v in agg.values())
    with open("profiles/launches_%s.txt" % R, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over: python bench.py --steps 1 --warmup 3 --n-steps 20000 --no-cpu-baseline\n")
        f.write("# (cold-cache, serialised: compare SHARES)\n%-60s %8s %12s %7s\n" % ("kernel", "launches", "total ms", "share"))
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-60s %8d %12.3f %6.1f%%\n" % (k[:60], n, ms, 100 * ms / tot))
import glob, shutil
for f in glob.glob("gpurun_out/traffic_*_%s.json" % R) + glob.glob("gpurun_out/traffic_writeback_%s.csv" % R) + glob.glob("gpurun_out/bench_%s_n*.json" % R) + glob.glob("gpurun_out/bench_ref_%s_n*.json" % R):
    shutil.copy(f, "profiles/")
for fn in ("bench_%s.json" % R, "bench_ref_%s.json" % R):
    p = os.path.join("gpurun_out", fn)
    if os.path.exists(p):
        open(os.path.join("profiles", fn), "w").write(open(p).read())
print(os.listdir("profiles"))
