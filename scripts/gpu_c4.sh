#!/bin/bash
# C4 development loop: parity tests of the triangulation kernels, then the bench for both engines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k triangulation 2>&1 | tail -15
for e in 1 2; do
  timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tri-engine $e > gpurun_out/c4_e$e.json 2> gpurun_out/c4_e$e.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c4_e$e.json").read().strip().splitlines()[-1])
    print("engine $e", d["ms_per_step"], "ms", d["value"], d["roofline"])
except Exception as ex:
    print("engine $e failed", ex); print(open("gpurun_out/c4_e$e.err").read()[-2000:])
PY
done
