#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== conv kernel tests (TMA-staged operand)"
  timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q 2>&1 | tail -12
  echo "== probe"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py
  echo "== resnet tests"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -x -q 2>&1 | tail -4
  echo "== bench"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-trainer-api --no-strong --no-e2e
  echo "== timeline A"
  DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_BWD=0 PROBE_EVENTS=110 timeout 120 python tools/conv_probe.py A_3x3_16_16@32
} > gpurun_out/r2_call9.log 2>&1
tail -3 gpurun_out/r2_call9.log
