#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  timeout 600 python tools/r2/diag_b128.py 128
  timeout 300 python tools/r2/diag_b128.py 16
} > gpurun_out/r2_call2.log 2>&1
tail -5 gpurun_out/r2_call2.log
