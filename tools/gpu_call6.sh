#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k wgrad > gpurun_out/gpu_tests6.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests6.log
tail -3 gpurun_out/gpu_tests6.log
for n in 3 2; do
  DPP_WG_NST=$n timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench6_n$n.json 2> gpurun_out/bench6_n$n.err
  python -c "import json;d=json.load(open('gpurun_out/bench6_n$n.json'));print('NST',$n,d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench6_n$n.err
  DPP_WG_NST=$n timeout 200 python tools/conv_probe.py > gpurun_out/probe6_n$n.log 2>&1
  grep -h "wgrad" gpurun_out/probe6_n$n.log
done
