#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
DPP_PDL=3 timeout 600 python -m pytest tests/test_gpu_resnet.py tests/test_gpu_conv_tc.py -m gpu -x -q > gpurun_out/gpu_tests_pdl.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests_pdl.log
tail -3 gpurun_out/gpu_tests_pdl.log
for m in 0 1 3; do
  DPP_PDL=$m timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_pdl$m.json 2> gpurun_out/bench_pdl$m.err
  python -c "import json;d=json.load(open('gpurun_out/bench_pdl$m.json'));print('PDL',$m,d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench_pdl$m.err
done
timeout 300 python tools/conv_probe.py > gpurun_out/probe2_k0.log 2>&1
grep -h "us" gpurun_out/probe2_k0.log | head -20
for sh in A_3x3_16_16@32 B_1x1_16_64@32+res; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_mn --launch-skip 2 --launch-count 1 \
     -f -o gpurun_out/ncu_wg_${sh%%_*} python tools/conv_probe.py $sh > gpurun_out/ncu_wg_${sh%%_*}.log 2>&1
done
