#!/bin/bash
# profiles of round 2: launch list of one step, --set full captures, sanitizer passes
mkdir -p gpurun_out
export PYTHONPATH=deep-prior-pp_b200:tests:$PYTHONPATH
echo "== launch list (ncu, one step in eager mode after warm-up)" > gpurun_out/r2_call13.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   -k regex:k_ -s 325 -c 175 --csv --log-file gpurun_out/r2_launches.csv \
   python tools/r2/one_step.py 4 >> gpurun_out/r2_call13.log 2>&1
echo "== full captures" >> gpurun_out/r2_call13.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -o gpurun_out/r2_ncu_conv_tc_A \
   python tools/conv_probe.py A_3x3_16_16@32 >> gpurun_out/r2_call13.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 12 -c 2 -o gpurun_out/r2_ncu_conv_tc_B \
   python tools/conv_probe.py B_1x1_16_64@32+res >> gpurun_out/r2_call13.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_group -s 4 -c 4 -o gpurun_out/r2_ncu_wgrad_group \
   python tools/r2/one_step.py 2 >> gpurun_out/r2_call13.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_augment -s 2 -c 1 -o gpurun_out/r2_ncu_augment \
   python tools/r2/one_step.py 2 >> gpurun_out/r2_call13.log 2>&1
echo "== compute-sanitizer memcheck (batch-4 train step)" >> gpurun_out/r2_call13.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/r2/one_step.py 1 4 > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_memcheck.txt >> gpurun_out/r2_call13.log
echo "== compute-sanitizer racecheck (batch-4 train step)" >> gpurun_out/r2_call13.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/r2/one_step.py 1 4 > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_racecheck.txt >> gpurun_out/r2_call13.log
tail -12 gpurun_out/r2_call13.log
