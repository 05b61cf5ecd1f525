#!/bin/bash
# trip 4: full GPU tests; same-box A/B of the coarse tile map and the float4 advect; ncu launch list + full capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_gpu_tests.log
tail -6 gpurun_out/r2d_gpu_tests.log
B="python bench.py --steps 8 --warmup 3 --no-extra --no-cpu-baseline"
timeout 600 $B > gpurun_out/r2d_bench_512.json 2> gpurun_out/r2d_bench.err
FS_NO_TILEMAP=1 timeout 600 $B > gpurun_out/r2d_bench_512_notile.json 2>> gpurun_out/r2d_bench.err
FS_NO_ADVECT_VEC4=1 timeout 600 $B > gpurun_out/r2d_bench_512_celladvect.json 2>> gpurun_out/r2d_bench.err
timeout 600 $B > gpurun_out/r2d_bench_512_again.json 2>> gpurun_out/r2d_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload 128 --no-cpu-baseline --no-extra --no-kernels > gpurun_out/r2d_bench_128.json 2>> gpurun_out/r2d_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'ms/step %.3f value %.3f e2e %.3f launches/step %.0f clocks %s'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['gpu_launches']/d['steps'], d['clocks']))
    print('   ', ' '.join('%s=%.4f'%(k['kernel'][:24],k['avg_launch_ms']) for k in d['roofline']['kernels']))
PY
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-extra --no-kernels"
timeout 600 $CMD > gpurun_out/r2d_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2d_launches.csv $CMD > gpurun_out/r2d_ncu.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/r2d_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:relax_vec4 -s 300 -c 2 -o gpurun_out/r2d_prof_relax $CMD > gpurun_out/r2d_ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 $CMD > gpurun_out/r2d_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_vec4 -s 2 -c 2 -o gpurun_out/r2d_prof_advect $CMD > gpurun_out/r2d_ncu_full2.log 2>&1
echo "ncu full2 exit $?"
