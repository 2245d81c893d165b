#!/bin/bash
# Evidence run: all GPU tests, smoke, bench (ours + reference arm), ncu launch list of the bench command,
# ncu --set full captures of the hot kernels.  Outputs under gpurun_out/p_*.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/p_smi.txt; nproc >> gpurun_out/p_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/p_smoke.log 2>&1; tail -2 gpurun_out/p_smoke.log
timeout 900 python bench.py --warmup 3 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; tail -2 gpurun_out/p_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/p_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --no-sweep --epochs 2 --train-iters 1 --cpu-seconds 0.2 > gpurun_out/p_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 1 -c 1 -f -o gpurun_out/p_step_8192_fused \
  python tools/profile_step.py 8192 128 stage_1 0 1 > gpurun_out/p_s1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 4 -c 1 -f -o gpurun_out/p_step_8192_single \
  python tools/profile_step.py 8192 8 stage_1 0 0 > gpurun_out/p_s2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 1 -c 1 -f -o gpurun_out/p_step_1M_fused \
  python tools/profile_step.py 1048576 16 stage_1 0 1 > gpurun_out/p_s3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 4 -c 1 -f -o gpurun_out/p_step_1M_single \
  python tools/profile_step.py 1048576 8 stage_1 0 0 > gpurun_out/p_s4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 1 -c 1 -f -o gpurun_out/p_step_house36_4096_fused \
  python tools/profile_step.py 4096 128 house 0 1 36 > gpurun_out/p_s5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_tc -s 1 -c 1 -f -o gpurun_out/p_tc_grad \
  python tools/tc_grad_check.py 131072 2 > gpurun_out/p_tc_grad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_infer|rtg_scan|adam_kernel|grad_reduce|mlp_grad_kernel' -s 20 -c 8 -f -o gpurun_out/p_ppo_misc \
  python tools/profile_ppo.py 8192 16 2 fp32 > gpurun_out/p_ppo_misc.log 2>&1
# summaries on the box (the .ncu-rep files with sources are ~16 MB each; gpurun returns at most 64 MiB)
for r in p_step_8192_fused p_step_8192_single p_step_1M_fused p_step_1M_single p_step_house36_4096_fused p_tc_grad p_ppo_misc; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r.summary.csv > gpurun_out/$r.summary.txt 2>&1
  python tools/ncu_source_lines.py gpurun_out/$r.ncu-rep 60 > gpurun_out/$r.lines.txt 2>&1
done
rm -f gpurun_out/p_step_8192_single.ncu-rep gpurun_out/p_step_1M_fused.ncu-rep gpurun_out/p_step_1M_single.ncu-rep gpurun_out/p_step_house36_4096_fused.ncu-rep gpurun_out/p_ppo_misc.ncu-rep
ls -la gpurun_out | grep " p_"; du -sh gpurun_out
