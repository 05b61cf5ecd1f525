#!/bin/bash
# programmatic dependent launch of the sweeps, A/B on one GPU; then the GPU test tier
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-kernels "$@" > gpurun_out/r2pdl_${name}.json 2> gpurun_out/r2pdl_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2pdl_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f value %.3f e2e_ms %.3f launches/step %.0f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps']))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
run thin_pdl "FS_PDL=1" --grid 512,512,64
run thin_nopdl "FS_PDL=0" --grid 512,512,64
run 128_pdl "FS_PDL=1" --workload 128
run 128_nopdl "FS_PDL=0" --workload 128
run 512_pdl "FS_PDL=1"
run 512_nopdl "FS_PDL=0"
run 32_pdl "FS_PDL=1" --workload 32
run 32_nopdl "FS_PDL=0" --workload 32
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2pdl_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2pdl_gpu_tests.log; tail -4 gpurun_out/r2pdl_gpu_tests.log
