#!/bin/bash
# round 2, first GPU trip: parity (incl. the fused two-stage sweeps), bench A/B, pair-kernel tuning, launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_info.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_gpu_tests.log
tail -8 gpurun_out/r2_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_512.json 2> gpurun_out/r2_bench_512.err; echo "bench exit $?"; cat gpurun_out/r2_bench_512.json; tail -3 gpurun_out/r2_bench_512.err
FS_NO_PAIR=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_512_nopair.json 2>> gpurun_out/r2_bench_512.err; cat gpurun_out/r2_bench_512_nopair.json
for rows in 8 12; do
  FS_PAIR_ROWS=$rows timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2_bench_512_rows$rows.json 2>> gpurun_out/r2_bench_512.err; cat gpurun_out/r2_bench_512_rows$rows.json
done
for zc in 16 32 128; do
  FS_PAIR_ZCHUNK=$zc timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2_bench_512_zc$zc.json 2>> gpurun_out/r2_bench_512.err; cat gpurun_out/r2_bench_512_zc$zc.json
done
timeout 600 python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_128.json 2>> gpurun_out/r2_bench_512.err; cat gpurun_out/r2_bench_128.json
timeout 600 python bench.py --steps 100 --warmup 5 --workload 32 --no-cpu-baseline --no-extra > gpurun_out/r2_bench_32.json 2>> gpurun_out/r2_bench_512.err; cat gpurun_out/r2_bench_32.json
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra --no-kernels > gpurun_out/r2_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra --no-kernels > gpurun_out/r2_ncu.log 2>&1
echo "ncu exit $?"
