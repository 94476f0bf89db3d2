#!/bin/bash
# ncu --set full capture of the C4 kernel (engine via $1, default 2) after a plain run exited 0
E=${1:-2}
KN=${2:-triangulation_stream_kernel}
mkdir -p gpurun_out
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --tri-engine $E"
$C4 > gpurun_out/plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KN -s 1 -c 1 -f -o gpurun_out/prof_tri_e$E $C4 > gpurun_out/ncu_c4_full.log 2>&1
tail -2 gpurun_out/plain_c4.log | cut -c1-300
