#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== conv / fc kernel tests"
  timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_convpool_fc.py -m gpu -x -q 2>&1 | tail -6
  echo "== per-layer timing"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py
  echo "== train-step parity"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -4
  echo "== bench"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong
} > gpurun_out/r2_call5.log 2>&1
tail -5 gpurun_out/r2_call5.log
