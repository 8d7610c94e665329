#!/bin/bash
# Run on the B200 box under gpurun: bench line + ncu launch list + ncu full captures -> gpurun_out/
set -x
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
CMD="python bench.py --steps 1 --warmup 3 --n-steps 20000 --no-cpu-baseline --no-config5"
$CMD > gpurun_out/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_list_$R.log 2>&1
$CMD > gpurun_out/plain2_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dram_kernel -s 3 -c 1 -o gpurun_out/prof_dram_$R $CMD > gpurun_out/ncu_dram_$R.log 2>&1
$CMD > gpurun_out/plain3_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ss_stream_kernel -s 3 -c 1 -o gpurun_out/prof_ss_$R $CMD > gpurun_out/ncu_ss_$R.log 2>&1
tail -n 2 gpurun_out/ncu_dram_$R.log; tail -n 2 gpurun_out/ncu_ss_$R.log
# big layout (BASELINE config 5's series length): second dram_kernel launch of the synthetic workload
python scripts/config5.py 1184 400 400 > gpurun_out/plain_c5_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dram_kernel -s 1 -c 1 -o gpurun_out/prof_c5_$R python scripts/config5.py 1184 400 400 > gpurun_out/ncu_c5_$R.log 2>&1
tail -n 2 gpurun_out/ncu_c5_$R.log
