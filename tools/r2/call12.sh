#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== bench default N=2"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -3
  echo "== test_gpu_dp"
  timeout 200 python -m pytest tests/test_gpu_dp.py -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/r2_call12.log 2>&1
tail -3 gpurun_out/r2_call12.log
