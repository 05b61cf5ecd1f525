#!/bin/bash
# Multi-GPU trip (gpurun --gpus N): slab parity tests, then bench at 1..N GPUs.  Every command is under
# `timeout` because a protocol bug would show up as kernels spinning on a flag.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.csv 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout -k 10 600 python -m pytest tests/test_slabs.py -m gpu -x -q > gpurun_out/slab_tests.log 2>&1; echo "slab tests exit $?" | tee -a gpurun_out/slab_tests.log
tail -12 gpurun_out/slab_tests.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout -k 10 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    fi
    echo "bench n=$n exit $?"; cut -c1-700 gpurun_out/scale_$n.json; tail -3 gpurun_out/scale_$n.err | cut -c1-400
  fi
done
