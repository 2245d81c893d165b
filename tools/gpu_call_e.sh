#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -15 gpurun_out/e_pytest.log
timeout 900 python bench.py --warmup 3 --precision bf16x3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -3 gpurun_out/e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['training']); print(d['other_configs']); print(d['roofline_sweep'])"
