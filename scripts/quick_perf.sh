#!/bin/bash
# Development aid (GPU box): the three quick throughput probes after a change of the forward model / the samplers.
python scripts/ss_bench.py 2>&1 | tail -2
python scripts/phases.py 299 40000 10000 2>&1 | tail -7
python scripts/warp_perf.py 8 4000 2000 0 2>&1 | tail -2
python scripts/warp_perf.py 64 2000 1000 0 2>&1 | tail -2
