#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  for s in A_3x3_16_16@32 C_1x1_64_16@32 E_3x3_64_64@8; do
    DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_BWD=0 PROBE_EVENTS=150 timeout 120 python tools/conv_probe.py $s
  done
} > gpurun_out/r2_call6.log 2>&1
tail -3 gpurun_out/r2_call6.log
