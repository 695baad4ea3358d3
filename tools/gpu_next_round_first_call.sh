#!/bin/bash
# First gpurun call of the NEXT round: everything round 1 wrote after its GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_next_round_first_call.sh'
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
L=gpurun_out/next_round_first_call.log
{
  echo "== 1. tests never run on hardware: reference-code fixture, opt-in 8-filter conv+pool kernels"
  timeout 200 python -m pytest tests/test_gpu_zz_reference_fixture.py -m gpu -q -s; echo "exit $?"
  DPP_TEST_CONVPOOL_FAST=1 timeout 200 python -m pytest tests/test_gpu_zz_convpool8.py -m gpu -q; echo "exit $?"
  echo "== 2. the nets that use those kernels, with the fast path on (must stay green before the default flips)"
  DPP_CONVPOOL_FAST=1 timeout 300 python -m pytest tests/test_gpu_scalenet.py tests/test_gpu_poseregnet.py tests/test_gpu_cascade.py -m gpu -q; echo "exit $?"
  echo "== 3. cascade bench, generic vs specialised conv+pool"
  DPP_CONVPOOL_FAST=0 timeout 200 python tools/bench_cascade.py --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline
  DPP_CONVPOOL_FAST=1 timeout 200 python tools/bench_cascade.py --batch 1024 --steps 10 --warmup 3 --no-cpu-baseline
  echo "== 4. headline bench incl. e2e_with_host_prep"
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline
} > $L 2>&1
tail -40 $L
