#!/bin/bash
# Round-2 evidence run: all GPU tests, smoke, bench (ours + reference arm), ncu launch list of the bench command,
# ncu --set full captures of the hot kernels, role / timeline profiles.  Outputs under gpurun_out/z_*.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/z_smi.txt; nproc >> gpurun_out/z_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
tail -4 gpurun_out/z_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/z_smoke.log 2>&1; tail -2 gpurun_out/z_smoke.log
timeout 900 python bench.py --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; tail -2 gpurun_out/z_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench_ref.json 2>&1
timeout 300 python tools/rollout_check.py > gpurun_out/z_rollout_ab.txt 2>&1; tail -6 gpurun_out/z_rollout_ab.txt
timeout 200 python tools/tc_ws_check.py 1048576 3 > gpurun_out/z_tc_ws_check.txt 2>&1; tail -7 gpurun_out/z_tc_ws_check.txt
timeout 200 python tools/tc_ws_profile.py --trace > gpurun_out/z_tc_ws_roles.txt 2>&1
timeout 200 python tools/tc_infer_profile.py > gpurun_out/z_infer_timeline.txt 2>&1
timeout 200 python tools/tc_infer_check.py > gpurun_out/z_infer_check.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/z_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --no-sweep --epochs 2 --train-iters 1 --cpu-seconds 0.2 > gpurun_out/z_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:navsim_step_kernel -s 1 -c 1 -f -o gpurun_out/z_step_8192_fused \
  python tools/profile_step.py 8192 128 stage_1 0 1 > gpurun_out/z_s1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_ws -s 1 -c 1 -f -o gpurun_out/z_tc_grad_ws \
  python tools/tc_grad_check.py 131072 2 > gpurun_out/z_tc_grad.log 2>&1
NAVPPO_ROLLOUT_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_infer_ws -s 3 -c 1 -f -o gpurun_out/z_tc_infer \
  python tools/tc_infer_profile.py 8192 > gpurun_out/z_tc_infer.log 2>&1
for r in z_step_8192_fused z_tc_grad_ws z_tc_infer; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r.summary.csv > gpurun_out/$r.summary.txt 2>&1
  python tools/ncu_source_lines.py gpurun_out/$r.ncu-rep 60 > gpurun_out/$r.lines.txt 2>&1
done
rm -f gpurun_out/z_step_8192_fused.ncu-rep gpurun_out/z_tc_infer.ncu-rep
ls -la gpurun_out | grep " z_"; du -sh gpurun_out
