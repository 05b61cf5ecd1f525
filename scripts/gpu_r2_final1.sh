#!/bin/bash
# end-of-round check on one GPU: smoke, the default bench line (with extra runs and CPU baseline), the reference arm
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -1 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench_512.json 2> gpurun_out/r2f_bench_512.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_512.json').read().strip().splitlines()[-1])
print('512 ms/step %.3f value %.3f e2e %.3f frame %.3f launches %d roofline %.3f'%(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('frame_value',0), d['gpu_launches'], d['roofline']['frac']))
for k in d['roofline'].get('kernels',[]): print('  %-45s %.4f ms  %.3f'%(k['kernel'],k['avg_launch_ms'],k['frac']))
print(json.dumps(d.get('extra'))[:1500])
print(json.dumps(d.get('cpu_baseline'))[:600])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_reference.json 2> gpurun_out/r2f_reference.err; echo "ref exit $?"; tail -c 700 gpurun_out/r2f_reference.json
