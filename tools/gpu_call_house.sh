#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/lane_sweep.py 4096 house 10
timeout 300 python tools/lane_sweep.py 4096 house 36
timeout 300 python tools/lane_sweep.py 32768 house 36
