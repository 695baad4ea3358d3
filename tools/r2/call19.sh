#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== bench default N=8"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -3
} > gpurun_out/r2_call19.log 2>&1
tail -2 gpurun_out/r2_call19.log | cut -c1-400
