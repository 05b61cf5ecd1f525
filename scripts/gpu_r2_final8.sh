#!/bin/bash
# 8 GPUs: the default schedule (single-launch exchange sweep + extended sweeps), without extended sweeps, and 1024^3
set -u
N=8
mkdir -p gpurun_out
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 10 --warmup 3 --no-extra --no-kernels "$@" > gpurun_out/r2f8_${name}.json 2> gpurun_out/r2f8_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2f8_${name}.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f value %.3f e2e_ms %.3f frame %.3f launches/step %.0f'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e'].get('frame_value',0), d['gpu_launches']/d['steps']), d.get('parity_check',{}).get('bit_exact'))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
run 512_ext "FS_EXTEND=1"
run 512_noext "FS_EXTEND=0"
run 1024_ext "FS_EXTEND=1" --workload 1024
run 512_ext_b "FS_EXTEND=1"
