#!/bin/bash
# Round-2 evidence for the trainer: all GPU tests, smoke, the training part of the bench, ncu --set full of the
# warp-specialised tcgen05 gradient kernel.  Outputs under gpurun_out/q_*.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/q_smoke.log 2>&1; tail -2 gpurun_out/q_smoke.log
timeout 900 python bench.py --warmup 3 --no-sweep --cpu-seconds 1 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_grad_ws -s 1 -c 1 -f -o gpurun_out/q_tc_grad_ws \
  python tools/tc_grad_check.py 131072 2 > gpurun_out/q_tc_grad.log 2>&1
python tools/ncu_summary.py gpurun_out/q_tc_grad_ws.ncu-rep gpurun_out/q_tc_grad_ws.summary.csv > gpurun_out/q_tc_grad_ws.summary.txt 2>&1
python tools/ncu_source_lines.py gpurun_out/q_tc_grad_ws.ncu-rep 60 > gpurun_out/q_tc_grad_ws.lines.txt 2>&1
tail -30 gpurun_out/q_tc_grad.log
ls -la gpurun_out | grep " q_"
