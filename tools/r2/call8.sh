#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  timeout 300 python tools/r2/step_breakdown.py 128
  timeout 300 python tools/r2/step_breakdown.py 64
} > gpurun_out/r2_call8.log 2>&1
tail -12 gpurun_out/r2_call8.log
