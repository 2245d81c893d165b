#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/lane_sweep.py 8192 stage_1 2>&1 | grep -E "lanes=( 4| 8|16|32)"
timeout 300 python tools/lane_sweep.py 4096 house 10 | grep -E "lanes=( 8|16|32)"
timeout 300 python tools/lane_sweep.py 4096 house 36 | grep -E "lanes=( 8|16|32)"
