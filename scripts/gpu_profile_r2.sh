#!/bin/bash
# Round-2 profile: plain runs first (must exit 0), then the ncu launch list of the default bench command and `--set full` captures of
# the dominant kernels: C5 engine 4 (tcgen05 cta_group::2), the bake-off engines 1 (LOP3+POPC) and 2 (mma.sync b1), C4 stream kernel
# in the compact-gather form.  Outputs under gpurun_out/ (summarised into profiles/ by scripts/ncu_summary.py).
mkdir -p gpurun_out
DEF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
C4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
SMALL="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary --nq 65536 --nd 1048576"
$DEF > gpurun_out/plain_def.json 2> gpurun_out/plain_def.err || exit 1
$C4 > gpurun_out/plain_c4.json 2> gpurun_out/plain_c4.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_default.csv $DEF --no-secondary > gpurun_out/ncu_def_list.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_c4_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn2_tc2_kernel -s 2 -c 1 -f -o gpurun_out/r02_knn2_tc2 $DEF --no-secondary > gpurun_out/ncu_c5_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triangulation_stream_kernel -s 4 -c 1 -f -o gpurun_out/r02_tri_stream $C4 > gpurun_out/ncu_c4_full.log 2>&1
for E in 1 2; do
  $SMALL --engine $E > gpurun_out/plain_e$E.json 2> gpurun_out/plain_e$E.err &&
  ncu --set full --clock-control none --import-source on -k regex:"knn2_(popc|mma_b1)_kernel" -s 1 -c 1 -f -o gpurun_out/r02_knn2_e$E $SMALL --engine $E > gpurun_out/ncu_e$E.log 2>&1
done
$SMALL --engine 4 > gpurun_out/plain_e4.json 2> gpurun_out/plain_e4.err
ls -la gpurun_out | tail -30
