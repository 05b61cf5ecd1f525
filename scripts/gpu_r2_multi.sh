#!/bin/bash
# multi-GPU trip: N = number of GPUs on the box (gpurun --gpus N).  Slab tests, then bench at N ranks with and without fused pairs.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo_$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 1200 python -m pytest tests/test_slabs.py tests/test_obstacles_and_sources.py -m gpu -x -q > gpurun_out/r2m_slab_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_slab_tests.log
  tail -6 gpurun_out/r2m_slab_tests.log
fi
run() { # name, extra env, bench args
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N "$@" > gpurun_out/r2m_${name}_$N.json 2> gpurun_out/r2m_${name}_$N.err
  echo "$name rc=$?"; tail -c 600 gpurun_out/r2m_${name}_$N.err | tail -2
}
run b512 "FS_X=0" --steps 10 --warmup 3
run b512_sep_push "FS_FUSED_PUSH=0" --steps 10 --warmup 3 --no-extra
run b512_weak "FS_X=0" --steps 5 --warmup 3 --scaling weak --no-extra --no-kernels
if [ "$N" = "8" ]; then
  run b1024 "FS_X=0" --steps 5 --warmup 3 --workload 1024 --no-extra
  run b1024rb "FS_X=0" --steps 5 --warmup 3 --workload 1024rb --no-extra
fi
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_*_$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'grid',d['config']['grid'],'ms/step %.3f value %.3f e2e %.3f frame %.3f launches/step %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['frame_value'],d['gpu_launches']/d['steps']), d.get('parity_check'))
    print('   ', ' '.join('%s=%.4f'%(k['kernel'][:24],k['avg_launch_ms']) for k in d['roofline']['kernels']))
    if d.get('extra'): print('   extra', {k:(v.get('ms_per_step'),v.get('value'),v.get('error')) for k,v in d['extra'].items()})
PY
