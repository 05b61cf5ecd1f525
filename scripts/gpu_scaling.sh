#!/bin/bash
# Scaling run on one 8-GPU box: bench.py at N = 1,2,4,8 (512^3 default workload) + 1024^3 Jacobi at N = 1 and 8.
set -u
mkdir -p gpurun_out
run() { # n workload tag
  local n=$1 w=$2 tag=$3
  if [ $n -eq 1 ]; then
    timeout -k 5 200 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  else
    timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2966$n bench.py --gpus $n --steps 5 --warmup 3 --workload $w > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  fi
  echo "$tag exit $?"; cut -c1-200 gpurun_out/$tag.json
}
run 8 512 scale_8
run 4 512 scale_4
run 2 512 scale_2
run 1 512 scale_1
run 8 1024 scale1024_8
run 1 1024 scale1024_1
