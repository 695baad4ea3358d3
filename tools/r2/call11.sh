#!/bin/bash
# 2 GPUs: data-parallel checks and benches
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== dp_check (2 ranks)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/dp_check.py 2>&1 | grep -v "^W\|^$" | tail -8
  echo "== bench default N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep -v "^W\|^$" | tail -3
  echo "== bench icvl512 N=2"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 2 --workload icvl512 --steps 10 --warmup 3 --no-e2e 2>&1 | grep -v "^W\|^$" | tail -3
  echo "== bench icvl512 N=2 syncbn"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29594 bench.py --gpus 2 --workload icvl512 --syncbn --steps 10 --warmup 3 --no-e2e 2>&1 | grep -v "^W\|^$" | tail -3
  echo "== trainer under torchrun (main_nyu script, 2 ranks, 2 epochs)"
  cd deep-prior-pp_b200 && DPP_NET=resnet DPP_EPOCHS=2 DPP_POSES=2e4 DPP_TRAIN=2048 DPP_VAL=256 DPP_VALFREQ=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29595 main_nyu_posereg_embedding.py 2>&1 | grep -v "^W\|^$" | tail -25
} > gpurun_out/r2_call11.log 2>&1
tail -3 gpurun_out/r2_call11.log
