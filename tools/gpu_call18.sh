#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_scalenet.py tests/test_gpu_trainer.py tests/test_gpu_poseregnet.py -m gpu -x -q -s > gpurun_out/gpu_tests18.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests18.log
tail -30 gpurun_out/gpu_tests18.log
