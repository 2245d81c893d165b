#!/bin/bash
# 2-GPU call: sharded update == single-GPU update (NCCL gradient all-reduce), bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/m_dist_check.log 2>&1
tail -8 gpurun_out/m_dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 3 --precision bf16x3 > gpurun_out/m_bench_2gpu.json 2> gpurun_out/m_bench_2gpu.err
tail -3 gpurun_out/m_bench_2gpu.err
python -c "
import json
for l in open('gpurun_out/m_bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e')}); print(d['training'])"
