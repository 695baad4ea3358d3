#!/bin/bash
# Final measurements of round 2b on one B200 (everything lands in gpurun_out/r2b_final_*; copied to profiles/ afterwards).
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
O=gpurun_out/r2b_final
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee ${O}_gpu_tests.log
echo "== bench (default)"; timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; tail -c 300 ${O}_bench.json; echo
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; tail -c 400 ${O}_bench_reference.json; echo
echo "== launch list"; tools/r2b/launches.sh r2b_final | head -30
echo "== probes"
timeout 200 python tools/conv_probe.py > ${O}_conv_probe.txt 2>&1; grep -c "fwd" ${O}_conv_probe.txt
timeout 100 python tools/fc_stem_probe.py > ${O}_fc_stem_probe.txt 2>&1; PROBE_COLD=1 timeout 100 python tools/fc_stem_probe.py >> ${O}_fc_stem_probe.txt 2>&1; cat ${O}_fc_stem_probe.txt
echo "== ncu --set full"
PROBE_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -f -o ${O}_ncu_conv_tc_A python tools/conv_probe.py A_3x3_16_16@32 > /dev/null 2>&1
PROBE_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -f -o ${O}_ncu_conv_tc_B python tools/conv_probe.py B_1x1_16_64@32+res > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_wgrad3|k_wgrad1|k_fc_stream|k_stem|k_augment|k_wgrad_group" -s 15 -c 15 -f -o ${O}_ncu_step python tools/r2/one_step.py 2 > /dev/null 2>&1
for r in conv_tc_A conv_tc_B step; do python tools/ncu_digest.py ${O}_ncu_$r.ncu-rep > ${O}_ncu_${r}_digest.txt 2>&1; done
grep -c "^kernel" ${O}_ncu_*_digest.txt
echo "== other workloads"
for w in icvl512 msra15 poseregnet cascade; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_$w.json 2> ${O}_bench_$w.err; python -c "
import json,sys
d=json.load(open('${O}_bench_$w.json')); print('$w', round(d['value']), d['ms_per_step'] if 'ms_per_step' in d else '', round(d['e2e']['value']))"; done
