#!/bin/bash
set -x
mkdir -p gpurun_out
for L in 1 16; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 1 -c 1 -f -o gpurun_out/c_step8192_l${L}_fused \
  python tools/profile_step.py 8192 128 stage_1 $L 1 > gpurun_out/c_l${L}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 4 -c 1 -f -o gpurun_out/c_step1M_l1_single \
  python tools/profile_step.py 1048576 8 stage_1 1 0 > gpurun_out/c_1M.log 2>&1
ls -la gpurun_out
