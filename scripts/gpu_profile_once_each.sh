#!/bin/bash
# ncu --set full of the single-frame kernels named in $1 (regex; default: the ones changed last) while scripts/gpu_once_each.py runs
# two passes; scripts/ncu_table.py keeps each kernel's last launch.  The report must stay well under gpurun's 64 MiB return limit.
K=${1:-'init_resolve_kernel|featvec_bow_build_kernel|bow_big_lists_kernel|bow_big_resolve_kernel|bow_match_kernel'}
mkdir -p gpurun_out
timeout 300 python scripts/gpu_once_each.py > gpurun_out/once_each.txt 2>&1 || { tail -5 gpurun_out/once_each.txt; exit 1; }
timeout 600 ncu --set full --clock-control none -k regex:"$K" -c 24 -f -o gpurun_out/prof_once_each python scripts/gpu_once_each.py > gpurun_out/ncu_once_each.log 2>&1
tail -2 gpurun_out/ncu_once_each.log; ls -la gpurun_out/
