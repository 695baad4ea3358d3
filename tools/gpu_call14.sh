#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_resnet.py -m gpu -x -q > gpurun_out/gpu_tests14.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests14.log
tail -3 gpurun_out/gpu_tests14.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
python -c "import json;d=json.load(open('gpurun_out/bench14.json'));print('bench14',d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench14.err
timeout 150 python tools/conv_probe.py 2>&1 | grep "wgrad"
