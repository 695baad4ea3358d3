#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_WG_TIMELINE=1 PROBE_EVENTS=3000 timeout 100 python tools/conv_probe.py B_1x1_16_64@32+res D_1x1_256_64@8 > gpurun_out/wg_timeline3.log 2>&1
grep -c "" gpurun_out/wg_timeline3.log
