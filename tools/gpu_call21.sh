#!/bin/bash
# last sanity call of round 1: smoke() + the dataset-stack test (everything else ran in calls 19/20)
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ; echo "smoke exit $?" ) > gpurun_out/gpu_smoke21.log 2>&1
timeout 50 python -m pytest tests/test_gpu_cascade.py -m gpu -q -k "dataset or joint" >> gpurun_out/gpu_smoke21.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_smoke21.log
tail -12 gpurun_out/gpu_smoke21.log
