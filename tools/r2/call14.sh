#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  for s in A_3x3_16_16@32 C_1x1_64_16@32; do
    DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_BWD=1 PROBE_WG_TIMELINE=1 PROBE_EVENTS=130 timeout 120 python tools/conv_probe.py $s
  done
  echo "== augment capture"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_augment -c 1 -o gpurun_out/r2_ncu_augment python tools/r2/one_step.py 1 128
} > gpurun_out/r2_call14.log 2>&1
tail -3 gpurun_out/r2_call14.log
