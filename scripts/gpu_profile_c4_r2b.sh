#!/bin/bash
# Round-2, after the post-role rework of the C4 kernel: plain run (must exit 0), ncu launch list and `--set full` capture of
# triangulation_stream_kernel in the compact-gather form, the pipeline timelines.  Outputs under gpurun_out/.
mkdir -p gpurun_out
C4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$C4 > gpurun_out/plain_c4b.json 2> gpurun_out/plain_c4b.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4b.csv $C4 > gpurun_out/ncu_c4b_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triangulation_stream_kernel -s 4 -c 1 -f -o gpurun_out/r02b_tri_stream $C4 > gpurun_out/ncu_c4b_full.log 2>&1
python scripts/ncu_summary.py gpurun_out/r02b_tri_stream.ncu-rep > gpurun_out/r02b_tri_stream_summary.txt 2>&1
for P in 4096 512; do MODE=compact PAIRS=$P python scripts/gpu_tri_timeline.py; done > gpurun_out/r02b_timeline.txt 2>&1
MODE=dense PAIRS=148 python scripts/gpu_tri_timeline.py >> gpurun_out/r02b_timeline.txt 2>&1
ls -la gpurun_out | tail -12
