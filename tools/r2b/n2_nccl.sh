#!/bin/bash
# two GPUs: step time of the default workload under different NCCL CTA budgets (the FC bucket's all-reduce shares the SMs
# with the backward chain)
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
CFGS=("default" "NCCL_MAX_CTAS=2" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=16 NCCL_MIN_CTAS=16" "NCCL_MAX_CTAS=24 NCCL_MIN_CTAS=24" "NCCL_MAX_CTAS=32 NCCL_MIN_CTAS=32")
if [ -n "$1" ]; then CFGS=("${CFGS[@]:$1}"); fi
for cfg in "${CFGS[@]}"; do
  if [ "$cfg" = "default" ]; then envs=""; else envs="$cfg"; fi
  r=$(env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e --no-strong --no-roofline --no-cost-check 2>/dev/null | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])")
  echo "$cfg: $r"
done
