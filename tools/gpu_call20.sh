#!/bin/bash
# pose sampling (f3) + trainer with the vectorised record worker on the B200
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 150 python -m pytest tests/test_gpu_poses.py tests/test_gpu_trainer.py tests/test_gpu_augment.py -m gpu -q -s > gpurun_out/gpu_tests20.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu_tests20.log
tail -15 gpurun_out/gpu_tests20.log
