#!/bin/bash
# ncu launch list (durations + DRAM bytes) of two eager training steps; usage: tools/r2b/launches.sh <tag>
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
tag=${1:-r2b}
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/r2/one_step.py 3 > gpurun_out/${tag}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv --steps 3 > gpurun_out/${tag}_launch_summary.txt 2>&1
head -40 gpurun_out/${tag}_launch_summary.txt
