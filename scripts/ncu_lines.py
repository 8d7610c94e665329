#!/usr/bin/env python
"""Summarise an ncu report by CUDA source line: share of stall samples and of executed instructions.
usage: ncu_lines.py report.ncu-rep [topN]   (run where ncu is installed; no GPU needed)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = ""; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "-") and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try: return float(d.get(k, 0) or 0)
            except ValueError: return 0.0
        lines.append((cur_file, int(r[0]), r[1], f("# Samples"), f("Instructions Executed"), d))
ts = sum(l[3] for l in lines) or 1; te = sum(l[4] for l in lines) or 1
print("total samples %d, warp instructions %d" % (ts, te))
stall_keys = [k for k in (lines[0][5].keys() if lines else []) if k.startswith("stall_")]
print("%-16s %5s %7s %7s  %s" % ("file", "line", "samp%", "inst%", "source"))
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-16s %5d %6.1f%% %6.1f%%  %s" % (l[0][:16], l[1], 100 * l[3] / ts, 100 * l[4] / te, l[2].strip()[:100]))
