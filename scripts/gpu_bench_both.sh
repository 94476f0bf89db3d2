set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 3000 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
