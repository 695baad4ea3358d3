#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
timeout 200 python -m pytest tests/test_gpu_trainer.py -m gpu -x -q -s > gpurun_out/gpu_tests17.log 2>&1; echo "pytest exit $?" >> gpurun_out/gpu_tests17.log
tail -25 gpurun_out/gpu_tests17.log
cd deep-prior-pp_b200 && DPP_TRAIN=1024 DPP_VAL=256 DPP_EPOCHS=3 DPP_VALFREQ=8 DPP_NET=resnet timeout 200 python main_nyu_posereg_embedding.py > ../gpurun_out/main_nyu.log 2>&1; echo "main rc=$?"; tail -8 ../gpurun_out/main_nyu.log
