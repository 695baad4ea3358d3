#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== train step parity with FC ReLU decisions pinned too"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s -k "train_step" 2>&1 | grep -v "^$" | tail -60
  echo "== per-layer timing (release build)"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py
  echo "== in-kernel timelines (prof build)"
  for s in A_3x3_16_16@32 C_1x1_64_16@32 B_1x1_16_64@32+res E_3x3_64_64@8 D_1x1_256_64@8; do
    DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_BWD=0 PROBE_EVENTS=260 timeout 120 python tools/conv_probe.py $s
  done
  echo "== wgrad timelines"
  for s in A_3x3_16_16@32 D_1x1_256_64@8; do
    DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_BWD=1 PROBE_WG_TIMELINE=1 PROBE_EVENTS=0 timeout 120 python tools/conv_probe.py $s
  done
} > gpurun_out/r2_call3.log 2>&1
tail -5 gpurun_out/r2_call3.log
