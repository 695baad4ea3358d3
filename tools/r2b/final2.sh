#!/bin/bash
# last measurements of the round (after the transform-loop trimming): bench (both arms), launch list, conv probe
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
O=gpurun_out/r2b_final
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee ${O}_gpu_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; python -c "
import json; d=json.load(open('${O}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['trainer_api']['value'], d['roofline']['frac'], d['cost_check']['rel'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; python -c "
import json; d=json.load(open('${O}_bench_reference.json')); print('reference', d['value'])"
tools/r2b/launches.sh r2b_final | head -12
timeout 200 python tools/conv_probe.py > ${O}_conv_probe.txt 2>&1; grep -c fwd ${O}_conv_probe.txt
