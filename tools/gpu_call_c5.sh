#!/bin/bash
# c5 (house map, table start / goal, beam sweep): FP32 instruction counts + DRAM bytes of one fused launch per beam
# count, and ncu --set full for B = 36
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
: > gpurun_out/c5_flops.txt
for b in 10 12 18 24 36; do
  echo "== beams $b" >> gpurun_out/c5_flops.txt
  timeout 300 ncu --metrics $M --clock-control none -k regex:navsim_step -s 3 -c 1 --csv python tools/profile_step.py 4096 128 house 0 1 $b small_house 2>/dev/null | grep -E "navsim_step|Metric" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' >> gpurun_out/c5_flops.txt
done
cat gpurun_out/c5_flops.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step -s 3 -c 1 -f -o gpurun_out/c5_house36 python tools/profile_step.py 4096 128 house 0 1 36 small_house > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/c5_house36.ncu-rep gpurun_out/c5_house36.summary.csv > gpurun_out/c5_house36.summary.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/c5_house36.ncu-rep 40 > gpurun_out/c5_house36.lines.txt 2>&1
rm -f gpurun_out/c5_house36.ncu-rep
tail -40 gpurun_out/c5_house36.summary.txt
