#!/bin/bash
# Profiling recipe for the count kernel (run on the GPU box through gpurun, one GPU).
#   gpurun --timeout 1200 -- 'bash profiles/run_ncu.sh sorted'
# Outputs land in gpurun_out/; the summaries kept for review are copied to profiles/ by hand.
set -u
ORDER=${1:-sorted}
READS=${READS:-10000000}   # READS=50000000 for the full cfg2 launch
CMD="python bench.py --reads $READS --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --order $ORDER"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$ORDER.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$ORDER.csv $CMD > gpurun_out/ncu_launches_$ORDER.log 2>&1
$CMD > gpurun_out/plain2_$ORDER.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sd_count_kernel -s 1 -c 1 -f -o gpurun_out/count_$ORDER $CMD > gpurun_out/ncu_full_$ORDER.log 2>&1
tail -2 gpurun_out/plain_$ORDER.log
