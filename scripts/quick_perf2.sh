#!/bin/bash
# Development aid (GPU box): config 2 at 100 000 steps + the two chain-per-warp probes, for the library in $TC_LIBTCMCMC
python scripts/ss_bench.py 2>&1 | tail -2 | head -1
python scripts/phases.py 299 100000 10000 2>&1 | grep "kernel\|cycles/step:"
python scripts/warp_perf.py 8 4000 2000 0 2>&1 | grep layout
python scripts/warp_perf.py 64 2000 1000 0 2>&1 | grep layout
