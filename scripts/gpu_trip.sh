#!/bin/bash
# One GPU trip: parity tests, smoke, bench, ncu launch list, ncu full capture of the hot sweep.
# Usage (under gpurun): bash scripts/gpu_trip.sh [tests|bench|ncu|all]
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
if [[ $what == all || $what == tests ]]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
  tail -15 gpurun_out/gpu_tests.log
  python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  python bench.py --steps 5 --warmup 3 --no-obstacle --no-cpu-baseline > gpurun_out/bench_noobst.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_noobst.json
  python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline > gpurun_out/bench_128.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_128.json
fi
if [[ $what == all || $what == ncu ]]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
  $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?"
  $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:relax_vec4 -s 60 -c 3 -o gpurun_out/prof_relax $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
fi
