#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
export DPP_WG_NST=2
for kn in 0 2 4 8 12; do
  echo "== WG_KNOBS $kn"
  DPP_WG_KNOBS=$kn timeout 120 python tools/conv_probe.py A_3x3_16_16@32 B_1x1_16_64@32+res C_1x1_64_16@32 2>&1 | grep wgrad
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python -c "import json;d=json.load(open('gpurun_out/bench7.json'));print('bench7',d['ms_per_step'],d['value'],d['e2e']['value'], d['gpu_launches'])" || tail -5 gpurun_out/bench7.err
