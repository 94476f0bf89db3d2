#!/bin/bash
# ncu --set full of the single-frame kernels (C1-C3, row a6) while scripts/gpu_latency.py runs them; plain run first.
mkdir -p gpurun_out
python scripts/gpu_latency.py > gpurun_out/latency.txt 2>&1 || { tail -5 gpurun_out/latency.txt; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k regex:'init_candidates_kernel|init_resolve_kernel|proj_candidates_kernel|proj_resolve_kernel|voc_transform_kernel|featvec_build_kernel|bow_match_kernel|grid_build_kernel' \
    --launch-skip 40 -c 16 -f -o gpurun_out/prof_single python scripts/gpu_latency.py > gpurun_out/ncu_single.log 2>&1
cat gpurun_out/latency.txt
