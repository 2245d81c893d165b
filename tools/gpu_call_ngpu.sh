#!/bin/bash
# N-GPU call (N = $1): sharded update == single-GPU update for every gradient-exchange mode, then the bench at N ranks
# with the NCCL, peer-memory and multimem exchanges.  Outputs under gpurun_out/n${N}_*.
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n${N}_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/n${N}_dist_check.log 2>&1
tail -14 gpurun_out/n${N}_dist_check.log
for ex in nccl peer nvls; do
NAVPPO_GRAD_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --precision bf16x3 --no-sweep > gpurun_out/n${N}_bench_$ex.json 2> gpurun_out/n${N}_bench_$ex.err
tail -3 gpurun_out/n${N}_bench_$ex.err
python -c "
import json
for l in open('gpurun_out/n${N}_bench_$ex.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$ex', {k:d[k] for k in ('value','n_gpus','ms_per_step')}); print(d['training'])"
done
