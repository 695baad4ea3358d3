#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== fused BN backward"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q -s -k "fused_bn or train_step" 2>&1 | grep -v "^$" | tail -25
  echo "== kernel tests"
  timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_convpool_fc.py -m gpu -x -q 2>&1 | tail -3
  echo "== probe"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py A_3x3_16_16@32 C_1x1_64_16@32 E_3x3_64_64@8
  echo "== bench fused / unfused"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
  DPP_FUSE_BN_BWD=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
} > gpurun_out/r2_call7.log 2>&1
tail -3 gpurun_out/r2_call7.log
