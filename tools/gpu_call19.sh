#!/bin/bash
# Cascade (f1) + metrics (f4) on the B200: parity tests, config-5 bench line, launch list with DRAM bytes.
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_cascade.py tests/test_gpu_poses.py tests/test_gpu_scalenet.py -m gpu -q -s > gpurun_out/gpu_tests19.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests19.log
tail -25 gpurun_out/gpu_tests19.log
timeout 150 python tools/bench_cascade.py --batch 1024 --steps 10 --warmup 3 > gpurun_out/cascade_bench_b1024.json 2> gpurun_out/cascade_bench_b1024.err
echo "bench b1024 exit $?"; tail -c 600 gpurun_out/cascade_bench_b1024.err; head -c 1500 gpurun_out/cascade_bench_b1024.json
if ! grep -q '"value"' gpurun_out/cascade_bench_b1024.json; then
  timeout 100 python tools/bench_cascade.py --batch 128 --steps 10 --warmup 3 > gpurun_out/cascade_bench_b128.json 2> gpurun_out/cascade_bench_b128.err
  echo "bench b128 exit $?"; tail -c 600 gpurun_out/cascade_bench_b128.err; head -c 1500 gpurun_out/cascade_bench_b128.json
fi
timeout 120 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 260 --csv \
  --log-file gpurun_out/cascade_launches.csv python tools/bench_cascade.py --batch 1024 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/cascade_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/cascade_launches.csv
