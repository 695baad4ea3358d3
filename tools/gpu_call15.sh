#!/bin/bash
# round-1 evidence run (final kernels): tests, both bench arms, launch list with DRAM bytes, ncu --set full digests
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python -c "import json;d=json.load(open('gpurun_out/bench_final.json'));print('final',d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'])"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 830 --launch-count 285 --csv \
   --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline --no-e2e > gpurun_out/ncu_bench_final.log 2>&1
grep -c k_conv_tc gpurun_out/launches_final.csv
for sh in A_3x3_16_16@32 B_1x1_16_64@32+res; do
  PROBE_EAGER=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_mn --launch-skip 5 --launch-count 1 \
     -f -o gpurun_out/ncu4_wg_${sh%%_*} python tools/conv_probe.py $sh > gpurun_out/ncu4_wg_${sh%%_*}.log 2>&1
done
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv,noheader
