#!/bin/bash
# one GPU, grid 512x512x64 (the per-GPU share of 512^3 on 8 GPUs): the no-exchange floor of a thin slab, z-chunk lengths
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-extra "$@" > gpurun_out/r2thin_${name}.json 2> gpurun_out/r2thin_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2thin_${name}.json').read().strip().splitlines()[-1])
    ks={k['kernel']:k['avg_launch_ms'] for k in d['roofline'].get('kernels',[])}
    print('${name}', 'ms/step %.3f value %.3f launches/step %.0f'%(d['ms_per_step'], d['value'], d['gpu_launches']/d['steps']), ' '.join('%s=%.4f'%(k.split('<')[-1][:12],v) for k,v in ks.items()))
except Exception as e:
    print('${name}', 'ERR', e)
PY
}
run default "FS_X=0" --grid 512,512,64
run z4 "FS_ZCHUNK=4" --grid 512,512,64
run z8 "FS_ZCHUNK=8" --grid 512,512,64
run z11 "FS_ZCHUNK=11" --grid 512,512,64
run z16 "FS_ZCHUNK=16" --grid 512,512,64
run z32 "FS_ZCHUNK=32" --grid 512,512,64
run nograph "FS_X=0" --grid 512,512,64 --no-graph
run g128 "FS_X=0" --grid 512,512,128
