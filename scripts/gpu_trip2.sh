#!/bin/bash
# GPU trip: parity tests, bench at several sizes, ncu full capture of the hot sweep.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -8 gpurun_out/gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 5 --warmup 3 --no-obstacle --no-cpu-baseline > gpurun_out/bench_noobst.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_noobst.json
python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline > gpurun_out/bench_128.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_128.json
python bench.py --steps 10 --warmup 3 --workload 256 --no-cpu-baseline > gpurun_out/bench_256.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_256.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
$CMD > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relax_vec4 -s 60 -c 2 -o gpurun_out/prof_relax $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
