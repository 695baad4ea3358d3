#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k wgrad > gpurun_out/gpu_tests13.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests13.log
tail -3 gpurun_out/gpu_tests13.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench13.json 2> gpurun_out/bench13.err
python -c "import json;d=json.load(open('gpurun_out/bench13.json'));print('bench13',d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench13.err
timeout 150 python tools/conv_probe.py 2>&1 | grep "us"
echo "== conv stats atomics ablated"
PROBE_BWD=0 DPP_TC_KNOBS=4 timeout 100 python tools/conv_probe.py 2>&1 | grep "us"
