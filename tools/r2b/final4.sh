#!/bin/bash
# final tree: kernel tests, default bench line, launch list
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
O=gpurun_out/r2b_final
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee ${O}_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['trainer_api']['value'], d['roofline']['frac'], d['cost_check']['rel'], d['cpu_baseline']['value'])"
tools/r2b/launches.sh r2b_final | head -8
