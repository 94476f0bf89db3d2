#!/bin/bash
# ncu evidence for the two bench workloads (run under gpurun; one GPU).  Each ncu run is preceded by the
# same command without ncu (B200_PROFILING.md).
set -x
mkdir -p gpurun_out
C5="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$C5 > gpurun_out/plain_c5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv $C5 > gpurun_out/ncu_c5_list.log 2>&1
$C5 > gpurun_out/plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2_tc_kernel -s 1 -c 1 -o gpurun_out/prof_knn2_tc $C5 > gpurun_out/ncu_c5_full.log 2>&1
$C4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_c4_list.log 2>&1
$C4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:triangulation_pairs_kernel -s 1 -c 1 -o gpurun_out/prof_tri $C4 > gpurun_out/ncu_c4_full.log 2>&1
ls -la gpurun_out
