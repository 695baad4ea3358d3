#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== full GPU suite"
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
  echo "== bench default"
  timeout 600 python bench.py --steps 20 --warmup 3
  echo "== bench msra15"
  timeout 300 python bench.py --workload msra15 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline
  echo "== bench icvl512 at N=1"
  timeout 300 python bench.py --workload icvl512 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
  echo "== reference arm"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1
} > gpurun_out/r2_call4.log 2>&1
tail -5 gpurun_out/r2_call4.log
