#!/bin/bash
# Profiling recipe for the lean count kernel (run on the GPU box through gpurun, one GPU).
#   gpurun --timeout 1200 -- 'bash profiles/run_ncu_lean.sh [tag]'
# Launch list of a short bench run, then one full-set capture of sd_lean_kernel (second launch), after the plain run exited 0.
set -u
TAG=${1:-lean}
READS=${READS:-50000000}
EXTRA=${EXTRA:-}
CMD="python bench.py --reads $READS --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $EXTRA"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:sd_lean_kernel' -s 1 -c 1 -f -o gpurun_out/count_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
