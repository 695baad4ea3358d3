#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== grouped wgrad"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q -s -k "grouped or fused_bn or train_step" 2>&1 | grep -v "^$" | tail -25
  echo "== conv tests"
  timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q 2>&1 | tail -3
  echo "== bench grouped / per-layer"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
  DPP_WGRAD_GROUP=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
  echo "== breakdown"
  timeout 300 python tools/r2/step_breakdown.py 128
} > gpurun_out/r2_call10.log 2>&1
tail -3 gpurun_out/r2_call10.log
