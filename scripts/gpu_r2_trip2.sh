#!/bin/bash
# quick trip: fused pair kernel v2 correctness + timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or vector_and or red_black or smooth_linsolve or config1 or project" > gpurun_out/r2b_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2b_bench_512.json 2> gpurun_out/r2b_bench.err; tail -2 gpurun_out/r2b_bench.err
for rows in 8 12; do
  FS_PAIR_ROWS=$rows timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2b_bench_512_rows$rows.json 2>> gpurun_out/r2b_bench.err
done
FS_L2_AHEAD=3 timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2b_bench_512_l2a3.json 2>> gpurun_out/r2b_bench.err
FS_L2_AHEAD=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2b_bench_512_l2a0.json 2>> gpurun_out/r2b_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_512*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'ms/step %.2f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k['kernel'][:28],k['avg_launch_ms']) for k in d['roofline']['kernels']))
PY
