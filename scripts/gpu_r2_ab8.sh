#!/bin/bash
set -u
N=8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slabs.py -m gpu -x -q > gpurun_out/r2ab8_slab_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ab8_slab_tests.log; tail -3 gpurun_out/r2ab8_slab_tests.log
run() { local name=$1; shift; local envs=$1; shift; local W=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-kernels --workload $W > gpurun_out/r2ab8_${name}.json 2> gpurun_out/r2ab8_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab8_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f value %.3f e2e_ms %.3f frame %.3f launches/step %.0f sweepJ %.4f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['frame_value'], d['gpu_launches']/d['steps'], d['roofline']['kernels'][0]['avg_launch_ms']), d.get('parity_check',{}).get('bit_exact'))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
run fused_1 "FS_X=0" 512
run sep_noprio_1 "FS_FUSED_PUSH=0 FS_HALO_NO_PRIORITY=1" 512
run fused_2 "FS_X=0" 512
run sep_prio_1 "FS_FUSED_PUSH=0" 512
run fused_p2 "FS_PUSH_PLANES=2" 512
run fused_1024 "FS_X=0" 1024
run sep_1024 "FS_FUSED_PUSH=0 FS_HALO_NO_PRIORITY=1" 1024
