#!/bin/bash
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --nq 65536 --nd 1048576 --engine 4"
$CMD > gpurun_out/plain_tc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 1 -c 1 -o gpurun_out/prof_knn2_tc2 $CMD > gpurun_out/ncu_tc2_full.log 2>&1
tail -1 gpurun_out/plain_tc2.log | cut -c1-200
