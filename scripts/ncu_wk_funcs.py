#!/usr/bin/env python
"""Aggregate an ncu report of dram_warp_kernel by source function (tc_warp.cuh / tc_device.cuh / tc_mcmc.cu): stall samples
and executed warp instructions.  usage: ncu_wk_funcs.py report.ncu-rep"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = ""; hdr = None; lines = []
def fl(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, int(r[0]), fl(d.get("# Samples", 0)), fl(d.get("Instructions Executed", 0))))
def marks(path):
    out = []
    pat = re.compile(r"^(?:template.*\n)?(?:__device__|__global__|static|inline|__host__)[^;{]*?\b(\w+)\s*\(", re.M)
    src = open(path).read()
    for m in pat.finditer(src):
        out.append((src.count("\n", 0, m.start()) + 1, m.group(1)))
    return out
M = {f: marks("transcriptioncycleinference_b200/csrc/" + f) for f in ("tc_warp.cuh", "tc_device.cuh", "tc_mcmc.cu")}
agg = {}
for f, l, sa, ie in lines:
    key = f
    if f in M:
        key = f + ":?"
        for a, n in M[f]:
            if a <= l: key = n
    agg.setdefault(key, [0, 0]); agg[key][0] += sa; agg[key][1] += ie
ts = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
print("total samples %d, warp instructions %.3e" % (ts, te))
for k, (sa, ie) in sorted(agg.items(), key=lambda x: -x[1][0]):
    if sa / ts > 0.002: print("%-28s samples %5.1f%%  inst %5.1f%%" % (k, 100 * sa / ts, 100 * ie / te))
