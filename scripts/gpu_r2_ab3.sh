#!/bin/bash
# extended sweeps (two per halo operation) vs one exchange per sweep, at N=2 on a thin-slab grid (512x512x128: 64 planes per
# GPU, like 512^3 on 8 GPUs) and on 512^3
set -u
N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slabs.py -m gpu -x -q > gpurun_out/r2ab3_slab_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ab3_slab_tests.log; tail -3 gpurun_out/r2ab3_slab_tests.log
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-kernels "$@" > gpurun_out/r2ab3_${name}.json 2> gpurun_out/r2ab3_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab3_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f value %.3f e2e_ms %.3f launches/step %.0f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps']), d.get('parity_check',{}).get('bit_exact'))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
for r in 1 2; do
run thin_ext_$r "FS_EXTEND=1" --grid 512,512,128
run thin_noext_$r "FS_EXTEND=0" --grid 512,512,128
done
run full_ext "FS_EXTEND=1"
run full_noext "FS_EXTEND=0"
run thin_ext_noobst "FS_EXTEND=1" --grid 512,512,128 --no-obstacle
