#!/bin/bash
# N-GPU trip 2: halo stream priority A/B, halo trace
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N "$@" > gpurun_out/r2n_${name}_$N.json 2> gpurun_out/r2n_${name}_$N.err
  echo "$name rc=$?"; tail -c 400 gpurun_out/r2n_${name}_$N.err | tail -1
}
run b512 "FS_X=0" --steps 20 --warmup 3 --no-extra --no-kernels
run b512_noprio "FS_HALO_NO_PRIORITY=1" --steps 20 --warmup 3 --no-extra --no-kernels
run b512_trace "FS_HALO_TRACE=gpurun_out/r2n_trace_$N" --steps 4 --warmup 3 --no-extra --no-kernels
run b1024 "FS_X=0" --steps 5 --warmup 3 --workload 1024 --no-extra --no-kernels
run b512_again "FS_X=0" --steps 20 --warmup 3 --no-extra --no-kernels
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_*_$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,'ERR',e); continue
    print(f, 'grid',d['config']['grid'],'ms/step %.3f value %.3f e2e %.3f frame %.3f launches/step %.0f'%(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['frame_value'],d['gpu_launches']/d['steps']), d.get('parity_check',{}).get('bit_exact'))
    print('   ', ' '.join('%s=%.4f'%(k['kernel'][:24],k['avg_launch_ms']) for k in d['roofline']['kernels']))
PY
ls -la gpurun_out/r2n_trace_* 2>/dev/null | head
