#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
for m in 2 3; do
  DPP_PDL=$m timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench3_pdl$m.json 2> gpurun_out/bench3_pdl$m.err
  python -c "import json;d=json.load(open('gpurun_out/bench3_pdl$m.json'));print('PDL',$m,d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench3_pdl$m.err
done
timeout 200 python tools/conv_probe.py > gpurun_out/probe3.log 2>&1
grep -h "us" gpurun_out/probe3.log | head -20
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv,noheader
