#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
{
  echo "== full GPU suite"
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== bench default"
  timeout 900 python bench.py --steps 20 --warmup 5
  echo "== bench reference arm"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1
  echo "== bench msra15 / icvl512 / poseregnet / cascade"
  timeout 300 python bench.py --workload msra15 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline
  timeout 300 python bench.py --workload icvl512 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
  timeout 300 python bench.py --workload poseregnet --steps 20 --warmup 3 --no-cpu-baseline
  timeout 300 python bench.py --workload cascade --steps 10 --warmup 3
  echo "== per-layer probe"
  PROBE_BWD=1 timeout 300 python tools/conv_probe.py
} > gpurun_out/r2_call18.log 2>&1
echo "== launch list" >> gpurun_out/r2_call18.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:k_ -s 200 -c 360 --csv --log-file gpurun_out/r2_final_launches.csv python tools/r2/one_step.py 4 >> gpurun_out/r2_call18.log 2>&1
tail -4 gpurun_out/r2_call18.log
