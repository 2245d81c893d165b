#!/bin/bash
# Round-2 extras: learning curves with the tensor-core rollout + update, compute-sanitizer over the new kernels.
set -x
mkdir -p gpurun_out
timeout 300 python tools/learn_curve.py 40 2048 512 50 bf16x3 stage_1 > gpurun_out/y_learning_curve_stage_1.jsonl 2> gpurun_out/y_lc1.err; tail -2 gpurun_out/y_learning_curve_stage_1.jsonl | cut -c1-400
timeout 300 python tools/learn_curve.py 40 2048 512 50 bf16x3 stage_2 > gpurun_out/y_learning_curve_stage_2.jsonl 2> gpurun_out/y_lc2.err; tail -1 gpurun_out/y_learning_curve_stage_2.jsonl | cut -c1-400
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python tools/profile_ppo.py 512 8 1 bf16x3 > gpurun_out/y_${tool}_ppo_tc.log 2>&1
  timeout 600 $CS --tool $tool --print-limit 20 python tools/profile_step.py 1500 12 stage_2 4 1 > gpurun_out/y_${tool}_step_l4.log 2>&1
done
NAVPPO_ROLLOUT_FUSED=0 timeout 900 $CS --tool memcheck --print-limit 20 python tools/profile_ppo.py 512 8 1 bf16x3 > gpurun_out/y_memcheck_ppo_tc_chained.log 2>&1
grep -H "ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/y_*.log
