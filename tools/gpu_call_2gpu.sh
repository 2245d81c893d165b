#!/bin/bash
# 2-GPU call: sharded update == single-GPU update for every gradient-exchange mode, bench at N=2 with the NCCL
# and the peer-memory exchange
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt
nvidia-smi topo -m >> gpurun_out/m_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/m_dist_check.log 2>&1
tail -14 gpurun_out/m_dist_check.log
for ex in nccl peer nvls; do
NAVPPO_GRAD_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --precision bf16x3 --no-sweep > gpurun_out/m_bench_2gpu_$ex.json 2> gpurun_out/m_bench_2gpu_$ex.err
tail -3 gpurun_out/m_bench_2gpu_$ex.err
python -c "
import json
for l in open('gpurun_out/m_bench_2gpu_$ex.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$ex', {k:d[k] for k in ('value','n_gpus','ms_per_step')}); print(d['training'])"
done
