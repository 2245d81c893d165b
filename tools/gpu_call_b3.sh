#!/bin/bash
set -x
timeout 1200 python -m pytest tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/lane_sweep.py 8192,1048576 stage_1 2>&1 | grep -E "lanes=( 1| 4| 8) "
