#!/bin/bash
# compute-sanitizer passes over the simulator and trainer kernels (small shapes)
set -x
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python tools/profile_step.py 1500 12 stage_2 4 1 > gpurun_out/z_${tool}_step_l4.log 2>&1
  timeout 600 $CS --tool $tool --print-limit 20 python tools/profile_step.py 700 12 house 16 1 36 > gpurun_out/z_${tool}_step_house.log 2>&1
  timeout 600 $CS --tool $tool --print-limit 20 python tools/profile_step.py 3000 6 stage_1 1 0 > gpurun_out/z_${tool}_step_l1.log 2>&1
  timeout 900 $CS --tool $tool --print-limit 20 python tools/profile_ppo.py 512 8 1 bf16x3 > gpurun_out/z_${tool}_ppo_tc.log 2>&1
  timeout 900 $CS --tool $tool --print-limit 20 python tools/profile_ppo.py 512 8 1 fp32 > gpurun_out/z_${tool}_ppo_fp32.log 2>&1
done
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY\|done" gpurun_out/z_*.log
