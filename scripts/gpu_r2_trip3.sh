#!/bin/bash
# trip 3: fused mirror + batched velocity sweeps: full GPU tests, bench at 512/128/32, launch list
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_gpu_tests.log
tail -6 gpurun_out/r2c_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench_512.json 2> gpurun_out/r2c_bench.err; tail -2 gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-obstacle --no-extra --no-cpu-baseline --no-kernels > gpurun_out/r2c_bench_512_noobst.json 2>> gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline --no-extra > gpurun_out/r2c_bench_128.json 2>> gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload 256 --no-cpu-baseline --no-extra --no-kernels > gpurun_out/r2c_bench_256.json 2>> gpurun_out/r2c_bench.err
timeout 600 python bench.py --steps 100 --warmup 5 --workload 32 --no-cpu-baseline --no-extra --no-kernels > gpurun_out/r2c_bench_32.json 2>> gpurun_out/r2c_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'ms/step %.3f value %.3f e2e %.3f frame %.3f launches/step %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['frame_value'],d['gpu_launches']/d['steps']))
    print('   ', ' '.join('%s=%.4f'%(k['kernel'][:24],k['avg_launch_ms']) for k in d['roofline']['kernels']))
    if d.get('extra'): print('   extra', {k:(v.get('ms_per_step'),v.get('value')) for k,v in d['extra'].items()})
    if d.get('cpu_baseline'): print('   cpu', d['cpu_baseline']['value'], d['cpu_baseline']['faithful']['value'])
PY
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra --no-kernels > gpurun_out/r2c_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra --no-kernels > gpurun_out/r2c_ncu.log 2>&1
echo "ncu exit $?"
