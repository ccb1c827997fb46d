#!/bin/bash
# Profiling recipe for the count kernel (run on the GPU box through gpurun, one GPU).
#   gpurun --timeout 1200 -- 'bash profiles/run_ncu.sh sorted [tag]'
# Outputs land in gpurun_out/; the summaries kept for review are copied to profiles/ by hand.
# The staged kernel is sd_count_kernel<W, SB, QB, true>; the plain one (<..., false>) follows it in every step.
set -u
ORDER=${1:-sorted}
TAG=${2:-$ORDER}
READS=${READS:-10000000}   # READS=50000000 for the full cfg2 launch
EXTRA=${EXTRA:-}
CMD="python bench.py --reads $READS --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --order $ORDER $EXTRA"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:sd_count_kernel.*bool.1' -s 1 -c 1 -f -o gpurun_out/count_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
