#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -6 gpurun_out/gpu_tests.log
python scripts/exp_data_dependence.py > gpurun_out/exp_data.log 2>&1; tail -25 gpurun_out/exp_data.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline > gpurun_out/bench_128.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_128.json
