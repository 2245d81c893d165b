#!/bin/bash
# 8-GPU bench line of the final tree (default gradient exchange = peer memory fused into Adam) + the parity check
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/f8_dist_check.log 2>&1
tail -3 gpurun_out/f8_dist_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-sweep > gpurun_out/f8_bench.json 2> gpurun_out/f8_bench.err
tail -2 gpurun_out/f8_bench.err
python -c "
import json
for l in open('gpurun_out/f8_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}); print('e2e', d['e2e']['value']); print(d['training'].get('bf16x3'))"
