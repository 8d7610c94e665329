#!/bin/bash
# Run on the B200 box under gpurun: bench lines + ncu launch list + ncu full captures + HBM traffic -> gpurun_out/
# usage: bash scripts/profile_gpu.sh <tag> [quick|traffic]     (quick: skip the bench lines, captures only; traffic: only the
#        two HBM-traffic captures bench.py's roofline.traffic quotes — rerun after any change of the kernel sources)
set -x
R=${1:-r2a}
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
LEGS="--no-cpu-baseline --no-config5 --no-config3 --no-writeback"
if [ "$2" = "traffic" ]; then
CMD2="python bench.py --steps 1 --warmup 1 $LEGS"
ncu --metrics $M --clock-control none -k regex:dram_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_dram_$R.csv $CMD2 > gpurun_out/ncu_tr_$R.log 2>&1 &&
python scripts/traffic_json.py gpurun_out/traffic_dram_$R.csv dram_kernel config2 200000 10000 $R
CMD4="python bench.py --workload config3 --steps 1 --warmup 1 $LEGS"
ncu --metrics $M --clock-control none -k regex:dram_warp_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_warp_$R.csv $CMD4 > gpurun_out/ncu_tr3_$R.log 2>&1 &&
python scripts/traffic_json.py gpurun_out/traffic_warp_$R.csv dram_warp_kernel config3 20000 10000 $R
exit 0
fi
if [ "$2" != "quick" ]; then
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || exit 1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
fi
LEGS="--no-cpu-baseline --no-config5 --no-config3 --no-writeback"
# --- config 2 (dram_kernel): launch list + full capture at 20 000 steps, HBM traffic at the benched 200 000 steps
CMD="python bench.py --steps 1 --warmup 3 --n-steps 20000 $LEGS"
$CMD > gpurun_out/plain_$R.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_list_$R.log 2>&1
$CMD > gpurun_out/plain2_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dram_kernel -s 3 -c 1 -f -o gpurun_out/prof_dram_$R $CMD > gpurun_out/ncu_dram_$R.log 2>&1
$CMD > gpurun_out/plain3_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ss_stream_kernel -s 3 -c 1 -f -o gpurun_out/prof_ss_$R $CMD > gpurun_out/ncu_ss_$R.log 2>&1
CMD2="python bench.py --steps 1 --warmup 1 $LEGS"
$CMD2 > gpurun_out/plain4_$R.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:dram_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_dram_$R.csv $CMD2 > gpurun_out/ncu_tr_$R.log 2>&1 &&
python scripts/traffic_json.py gpurun_out/traffic_dram_$R.csv dram_kernel config2 200000 10000 $R
# --- config 3 (dram_warp_kernel): full capture at 4 000 steps, HBM traffic at the benched 20 000 steps
CMD3="python bench.py --workload config3 --steps 1 --warmup 1 --n-steps 4000 --n-burn 1000 $LEGS"
$CMD3 > gpurun_out/plain5_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dram_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_warp_$R $CMD3 > gpurun_out/ncu_warp_$R.log 2>&1
CMD4="python bench.py --workload config3 --steps 1 --warmup 1 $LEGS"
$CMD4 > gpurun_out/plain6_$R.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:dram_warp_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_warp_$R.csv $CMD4 > gpurun_out/ncu_tr3_$R.log 2>&1 &&
python scripts/traffic_json.py gpurun_out/traffic_warp_$R.csv dram_warp_kernel config3 20000 10000 $R
# --- raw-chain write-back (config 2, 20 000 steps, store_chain = 1): HBM write rate of that launch
CMD5="python scripts/writeback_run.py"
$CMD5 > gpurun_out/plain7_$R.log 2>&1 &&
ncu --metrics $M,dram__bytes_write.sum.per_second --clock-control none -k regex:dram_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_writeback_$R.csv $CMD5 > gpurun_out/ncu_wb_$R.log 2>&1
tail -n 2 gpurun_out/ncu_dram_$R.log; tail -n 2 gpurun_out/ncu_ss_$R.log; tail -n 2 gpurun_out/ncu_warp_$R.log
# --- big layout (BASELINE config 5's series length): second dram_kernel launch of the synthetic workload
python scripts/config5.py 1184 400 400 > gpurun_out/plain_c5_$R.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dram_kernel -s 1 -c 1 -f -o gpurun_out/prof_c5_$R python scripts/config5.py 1184 400 400 > gpurun_out/ncu_c5_$R.log 2>&1
tail -n 2 gpurun_out/ncu_c5_$R.log
