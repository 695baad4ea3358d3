#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py > gpurun_out/dp_check.log 2>&1
tail -4 gpurun_out/dp_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1200 gpurun_out/bench_n2.json; tail -c 400 gpurun_out/bench_n2.err
DPP_EARLY_ALLREDUCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-roofline > gpurun_out/bench_n2_late.json 2> gpurun_out/bench_n2_late.err
tail -c 700 gpurun_out/bench_n2_late.json
