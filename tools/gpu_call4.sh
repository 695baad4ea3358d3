#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 400 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_resnet.py -m gpu -x -q > gpurun_out/gpu_tests4.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests4.log
tail -4 gpurun_out/gpu_tests4.log
for kn in 0 2; do
  DPP_TC_KNOBS=$kn timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench4_k$kn.json 2> gpurun_out/bench4_k$kn.err
  python -c "import json;d=json.load(open('gpurun_out/bench4_k$kn.json'));print('KNOB',$kn,d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench4_k$kn.err
done
timeout 200 python tools/conv_probe.py > gpurun_out/probe4_k0.log 2>&1
grep -h "us" gpurun_out/probe4_k0.log | head -20
PROBE_BWD=0 DPP_TC_KNOBS=2 timeout 200 python tools/conv_probe.py > gpurun_out/probe4_k2.log 2>&1
grep -h "us" gpurun_out/probe4_k2.log | head -20
