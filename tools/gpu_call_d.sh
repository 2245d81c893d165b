#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --precision bf16x3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -3 gpurun_out/d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/d_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print(d['roofline_sweep']); print(d['training'])"
