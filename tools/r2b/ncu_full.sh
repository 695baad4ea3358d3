#!/bin/bash
# ncu --set full of the round-2b kernels inside the second of two eager training steps; usage: tools/r2b/ncu_full.sh <tag> [regex] [skip] [count]
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
tag=${1:-r2b}; rx=${2:-"k_wgrad3|k_fc_stream|k_stem"}; skip=${3:-10}; cnt=${4:-10}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/${tag}_ncu_full \
    python tools/r2/one_step.py 2 > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_digest.py gpurun_out/${tag}_ncu_full.ncu-rep > gpurun_out/${tag}_ncu_full_digest.txt 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
