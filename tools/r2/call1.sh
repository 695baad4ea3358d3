#!/bin/bash
# round 2, call 1: the new batch-128 parity tests + everything round 1 left unverified + baseline bench
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
L=gpurun_out/r2_call1.log
{
  echo "== 1. new parity tests at the benchmarked shapes"
  timeout 900 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x -k "bench_shapes" -s 2>&1 | tail -50; echo "exit $?"
  timeout 600 python -m pytest tests/test_gpu_resnet.py -m gpu -q -s -k "train_step" 2>&1 | tail -30; echo "exit $?"
  echo "== 2. never run on hardware"
  timeout 200 python -m pytest tests/test_gpu_zz_reference_fixture.py -m gpu -q 2>&1 | tail -5
  DPP_TEST_CONVPOOL_FAST=1 timeout 200 python -m pytest tests/test_gpu_zz_convpool8.py -m gpu -q 2>&1 | tail -5
  DPP_CONVPOOL_FAST=1 timeout 300 python -m pytest tests/test_gpu_scalenet.py tests/test_gpu_poseregnet.py tests/test_gpu_cascade.py -m gpu -q 2>&1 | tail -5
  echo "== 3. cascade bench, generic vs specialised conv+pool"
  DPP_CONVPOOL_FAST=0 timeout 200 python tools/bench_cascade.py --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline
  DPP_CONVPOOL_FAST=1 timeout 200 python tools/bench_cascade.py --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline
  echo "== 4. headline bench"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
} > $L 2>&1
tail -60 $L
