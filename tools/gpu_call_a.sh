#!/bin/bash
# GPU call: tests, bench (bf16x3 training), launch list of a training iteration, full captures of the PPO kernels
set -x
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/a_smi.txt
nproc >> gpurun_out/a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16x3 > gpurun_out/a_bench_bf16x3.json 2> gpurun_out/a_bench_bf16x3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/a_launches_train.csv \
  python bench.py --steps 1 --warmup 1 --no-sweep --horizon 8 --epochs 2 --train-iters 1 --precision bf16x3 --cpu-seconds 0.2 --agents 8192 > gpurun_out/a_launches_train.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_tc -s 1 -c 1 -f -o gpurun_out/a_tc_grad \
  python tools/tc_grad_check.py 131072 2 > gpurun_out/a_tc_grad.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_infer|rtg_scan|adam_kernel|grad_reduce' -s 20 -c 8 -f -o gpurun_out/a_ppo_misc \
  python tools/profile_ppo.py 8192 16 2 bf16x3 > gpurun_out/a_ppo_misc.log 2>&1
ls -la gpurun_out
