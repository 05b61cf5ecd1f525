#!/bin/bash
# A/B/C/D of the exchange variants at N ranks, two rounds each, interleaved
set -u
N=${1:-2}; W=${2:-512}
mkdir -p gpurun_out
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 20 --warmup 3 --no-extra --no-kernels --workload $W > gpurun_out/r2ab_${name}_$N.json 2> gpurun_out/r2ab_${name}_$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2ab_${name}_$N.json').read().strip().splitlines()[-1])
    print('${name}', 'ms/step %.3f e2e_ms %.3f frame %.3f launches/step %.0f sweepJ %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['frame_value'], d['gpu_launches']/d['steps'], d['roofline']['kernels'][0]['avg_launch_ms']), d.get('parity_check',{}).get('bit_exact'))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
for round in 1 2; do
  run fused_prio_$round "FS_X=0"
  run sep_prio_$round "FS_FUSED_PUSH=0"
  run sep_noprio_$round "FS_FUSED_PUSH=0 FS_HALO_NO_PRIORITY=1"
  run fused_noprio_$round "FS_HALO_NO_PRIORITY=1"
done
