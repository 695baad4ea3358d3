#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== dp_check (2 ranks) incl. peer-memory SyncBN"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/dp_check.py 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -12
  echo "== bench icvl512 N=2 syncbn p2p"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29594 bench.py --gpus 2 --workload icvl512 --syncbn p2p --steps 10 --warmup 3 --no-e2e 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -3
  echo "== bench icvl512 N=2 syncbn nccl"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29595 bench.py --gpus 2 --workload icvl512 --syncbn nccl --steps 10 --warmup 3 --no-e2e 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -3
  echo "== bench icvl512 N=2"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29596 bench.py --gpus 2 --workload icvl512 --steps 10 --warmup 3 --no-e2e 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" | tail -3
} > gpurun_out/r2_call17.log 2>&1
tail -3 gpurun_out/r2_call17.log
