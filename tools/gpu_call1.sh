#!/bin/bash
# round-1 GPU call: full GPU tests, bench (both arms), kernel probes with tuning knobs, ncu captures
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 600 gpurun_out/bench_a.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for kn in 0 1; do
  DPP_TC_KNOBS=$kn DPP_WG_KNOBS=$kn timeout 300 python tools/conv_probe.py > gpurun_out/probe_k$kn.log 2>&1
done
grep -h "us" gpurun_out/probe_k0.log | head -20
for sh in A_3x3_16_16@32 B_1x1_16_64@32+res E_3x3_64_64@8; do
  PROBE_BWD=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc --launch-skip 4 --launch-count 1 \
     -f -o gpurun_out/ncu_${sh%%_*} python tools/conv_probe.py $sh > gpurun_out/ncu_${sh%%_*}.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_a.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_a.log 2>&1
ls -la gpurun_out | tail -20
