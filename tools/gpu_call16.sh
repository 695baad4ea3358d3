#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
t0=$(date +%s)
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_check.py > gpurun_out/dp_check.log 2>&1
echo "dp_check rc=$? after $(( $(date +%s) - t0 )) s"; grep "dp_check\|DP_CHECK\|Error" gpurun_out/dp_check.log | tail -4
t0=$(date +%s)
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$? after $(( $(date +%s) - t0 )) s"
grep '^{' gpurun_out/bench_n2.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('n2',d['ms_per_step'],d['value'],d['e2e']['value'])"
