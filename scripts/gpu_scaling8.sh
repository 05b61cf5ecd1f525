#!/bin/bash
# Final-scaling subset on one 8-GPU box: slab parity tests, then N = 8 and 4 at 512^3 and N = 8 at 1024^3.
set -u
mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_slabs.py -m gpu -x -q > gpurun_out/slab_tests.log 2>&1; echo "slab tests exit $?"; tail -2 gpurun_out/slab_tests.log
run() { local n=$1 w=$2 tag=$3
  timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2967$n bench.py --gpus $n --steps 5 --warmup 3 --workload $w > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  echo "$tag exit $?"; cut -c1-160 gpurun_out/$tag.json; }
run 8 512 scale_8
run 4 512 scale_4
run 8 1024 scale1024_8
