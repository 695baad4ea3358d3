#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -m gpu -x -q -k wgrad > gpurun_out/gpu_tests5.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests5.log
tail -3 gpurun_out/gpu_tests5.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python -c "import json;d=json.load(open('gpurun_out/bench5.json'));print('bench5',d['ms_per_step'],d['value'],d['e2e']['value'])" || tail -5 gpurun_out/bench5.err
timeout 200 python tools/conv_probe.py A_3x3_16_16@32 B_1x1_16_64@32+res C_1x1_64_16@32 H_3x3_32_32@16 D_1x1_256_64@8 > gpurun_out/probe5.log 2>&1
grep -h "wgrad" gpurun_out/probe5.log
for sh in A_3x3_16_16@32 B_1x1_16_64@32+res; do
  PROBE_EAGER=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_mn --launch-skip 5 --launch-count 1 \
     -f -o gpurun_out/ncu_wg2_${sh%%_*} python tools/conv_probe.py $sh > gpurun_out/ncu_wg2_${sh%%_*}.log 2>&1
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv,noheader
