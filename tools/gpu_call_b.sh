#!/bin/bash
# GPU call: parity tests of the lane-cooperative step kernel + lane sweep timings
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_gpu.py -m gpu -x -q > gpurun_out/b_pytest_env.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest_env.log
tail -5 gpurun_out/b_pytest_env.log
timeout 600 python tools/lane_sweep.py 8192,16384,65536,1048576 stage_1 > gpurun_out/b_lane_sweep.log 2>&1
timeout 300 python tools/lane_sweep.py 4096 house 10 >> gpurun_out/b_lane_sweep.log 2>&1
timeout 300 python tools/lane_sweep.py 4096 house 36 >> gpurun_out/b_lane_sweep.log 2>&1
timeout 300 python tools/lane_sweep.py 32768 house 36 >> gpurun_out/b_lane_sweep.log 2>&1
timeout 300 python tools/lane_sweep.py 16384 stage_2 10 >> gpurun_out/b_lane_sweep.log 2>&1
cat gpurun_out/b_lane_sweep.log
