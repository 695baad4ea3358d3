#!/bin/bash
# two GPUs: data-parallel checks (tools/dp_check.py = tests/test_gpu_dp.py) and the default bench at N = 2
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
O=gpurun_out/r2b_final
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/dp_check.py 2>&1 | grep -v "^W\|^$\|^\*\*\*\|OMP_NUM" > ${O}_dp_check_n2.txt; tail -12 ${O}_dp_check_n2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 5 2> ${O}_bench_n2.err | grep "^{" > ${O}_bench_n2.json; python -c "
import json; d=json.load(open('${O}_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['strong']['ms_per_step'], d['strong'].get('syncbn_ms_per_step'), d['strong'].get('syncbn_p2p_ms_per_step'))"
tail -3 ${O}_bench_n2.err
