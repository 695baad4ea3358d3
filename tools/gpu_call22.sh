#!/bin/bash
# round-1 last call: the headline bench line with the vectorised record preparation (no CPU baseline: budget)
mkdir -p gpurun_out
timeout 70 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench22.json 2> gpurun_out/bench22.err
echo "bench exit $?"; tail -c 400 gpurun_out/bench22.err; head -c 700 gpurun_out/bench22.json
