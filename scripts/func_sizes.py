#!/usr/bin/env python
"""Development aid: SASS size (bytes) of every device function of a built libtcmcmc.so, from the cubin's symbol table."""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "transcriptioncycleinference_b200/libtcmcmc.so"
txt = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
rows = []
for l in txt.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+|\d+)\s+(0x[0-9a-f]+|\d+)\s+0x(2|22|12)\s+\S+\s+\S+\s+(\S+)", l)
    if m:
        rows.append((int(m.group(2), 0), m.group(4)))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for sz, name in sorted(rows):
    if flt in name:
        name = re.sub(r"^\$_Z\d+(\w+?)7RunArgs\$|^\$_Z\d+(\w+?)6SsArgs\$", lambda m: (m.group(1) or m.group(2)) + "::", name)
        print("%7d  %s" % (sz, name[:140]))
