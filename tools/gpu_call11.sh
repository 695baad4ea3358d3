#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k wgrad > gpurun_out/gpu_tests11.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests11.log
tail -3 gpurun_out/gpu_tests11.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench11.json 2> gpurun_out/bench11.err
python -c "import json;d=json.load(open('gpurun_out/bench11.json'));print('bench11',d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench11.err
PROBE_BWD=1 timeout 150 python tools/conv_probe.py 2>&1 | grep wgrad
DPP_LIB=deep-prior-pp_b200/csrc/libdpp_b200_prof.so PROBE_WG_TIMELINE=1 PROBE_EVENTS=100 timeout 100 python tools/conv_probe.py B_1x1_16_64@32+res A_3x3_16_16@32 > gpurun_out/wg_timeline2.log 2>&1
