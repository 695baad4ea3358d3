#!/bin/bash
# evidence on the final build: sanitizer passes, ncu --set full of k_conv_tc on two layer shapes
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
O=gpurun_out/r2b_final
tools/r2b/sanitize.sh r2b
PROBE_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -f -o ${O}_ncu_conv_tc_A python tools/conv_probe.py A_3x3_16_16@32 > /dev/null 2>&1
PROBE_EAGER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -f -o ${O}_ncu_conv_tc_B python tools/conv_probe.py B_1x1_16_64@32+res > /dev/null 2>&1
for r in conv_tc_A conv_tc_B; do python tools/ncu_digest.py ${O}_ncu_$r.ncu-rep > ${O}_ncu_${r}_digest.txt 2>&1; grep "kernel\|duration\|issue_active\|pipe_tc_cycles\|bank_conflicts\|wavefronts_mem_shared.sum \|dram__bytes_read" ${O}_ncu_${r}_digest.txt | head -8; done
