#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== conv + resnet tests"
  timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -6
  echo "== bench"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
  echo "== bench poseregnet"
  timeout 600 python bench.py --workload poseregnet --steps 20 --warmup 3 --no-cpu-baseline
} > gpurun_out/r2_call16.log 2>&1
tail -3 gpurun_out/r2_call16.log
