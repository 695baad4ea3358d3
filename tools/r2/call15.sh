#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== conv kernel tests (TMA-staged wgrad)"
  timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q 2>&1 | tail -12
  echo "== resnet tests"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -4
  echo "== probe"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py A_3x3_16_16@32 B_1x1_16_64@32+res C_1x1_64_16@32 D_1x1_256_64@8 E_3x3_64_64@8 H_3x3_32_32@16
  echo "== bench"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
} > gpurun_out/r2_call15.log 2>&1
tail -3 gpurun_out/r2_call15.log
