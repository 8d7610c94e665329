#!/usr/bin/env python
"""Opcode histogram per kernel of libtcmcmc.so (cuobjdump -sass; no GPU needed) -> profiles/sass_hist_<tag>.txt.
The mnemonics that prove the Blackwell-native shape of an FP64 path: DMMA (mma.sync.m8n8k4.f64: tcgen05/TMEM has no f64
kind, so this is the only FP64 tensor path), DFMA/DADD/DMUL, LDGSTS (cp.async), UBLKCP (cp.async.bulk = TMA 1-D),
SYNCS (mbarrier), no HMMA/UTMALDG (nothing here is a dense low-precision GEMM or a 2-D tile)."""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "transcriptioncycleinference_b200", "libtcmcmc.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
hist = collections.OrderedDict(); cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEY = ["DMMA", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "F2F", "LDGSTS", "UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL",
       "BAR", "WARPSYNC", "ATOMS", "ATOMG", "HMMA", "IMMA", "UTMALDG", "UTCHMMA", "CCTL", "NANOSLEEP"]
out = ["# cuobjdump -sass transcriptioncycleinference_b200/libtcmcmc.so (sm_100a), opcode counts per kernel (device functions included)", ""]
for k, c in hist.items():
    tot = sum(c.values())
    out.append("%s: %d instructions (%.1f KB)" % (k, tot, tot * 16 / 1024))
    out.append("  " + "  ".join("%s %d" % (n, c[n]) for n in KEY if c[n]))
    out.append("  top: " + "  ".join("%s %d" % kv for kv in c.most_common(12)))
    out.append("")
p = os.path.join(ROOT, "profiles", "sass_hist_%s.txt" % tag)
open(p, "w").write("\n".join(out))
print("\n".join(out[:40]))
