#!/bin/bash
# Round profile: plain bench runs first (must exit 0), then the ncu launch lists of the same commands and one
# `--set full` capture of the dominant kernel of each workload.  Outputs under gpurun_out/ (summarised into profiles/).
mkdir -p gpurun_out
C5="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
C4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err || exit 1
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv $C5 > gpurun_out/ncu_c5_list.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_c4_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 2 -c 1 -f -o gpurun_out/prof_knn2_tc2 $C5 > gpurun_out/ncu_c5_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triangulation_stream_kernel -s 2 -c 1 -f -o gpurun_out/prof_tri_stream $C4 > gpurun_out/ncu_c4_full.log 2>&1
tail -c 600 gpurun_out/bench_c5.json; echo; tail -c 900 gpurun_out/bench_c4.json
