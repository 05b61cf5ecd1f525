#!/bin/bash
# single-launch exchange sweep (ends, push CTAs, middle in one launch) vs fork/join with a separate push kernel; N=2,
# thin-slab grid 512x512x128 (64 planes per GPU, like 512^3 on 8 GPUs) and 512^3
set -u
N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slabs.py -m gpu -x -q > gpurun_out/r2ab4_slab_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ab4_slab_tests.log; tail -3 gpurun_out/r2ab4_slab_tests.log
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-kernels "$@" > gpurun_out/r2ab4_${name}.json 2> gpurun_out/r2ab4_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab4_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f value %.3f e2e_ms %.3f launches/step %.0f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches']/d['steps']), d.get('parity_check',{}).get('bit_exact'))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
run thin_one_ext "FS_XCHG_IN_SWEEP=1" --grid 512,512,128
run thin_one_noext "FS_XCHG_IN_SWEEP=1 FS_EXTEND=0" --grid 512,512,128
run thin_fork_ext "FS_XCHG_IN_SWEEP=0" --grid 512,512,128
run thin_one_ext_p8 "FS_PUSH_CTAS=8" --grid 512,512,128
run thin_one_ext_p48 "FS_PUSH_CTAS=48" --grid 512,512,128
run thin_one_ext_b "FS_XCHG_IN_SWEEP=1" --grid 512,512,128
run full_one_ext "FS_XCHG_IN_SWEEP=1"
run full_one_noext "FS_XCHG_IN_SWEEP=1 FS_EXTEND=0"
