#!/bin/bash
set -x
mkdir -p gpurun_out
NG=${1:-8}
nvidia-smi -L > gpurun_out/s_smi.txt; nproc >> gpurun_out/s_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 100 --warmup 3 > gpurun_out/s_bench_${NG}gpu.json 2> gpurun_out/s_bench_${NG}gpu.err
tail -3 gpurun_out/s_bench_${NG}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --impl reference --steps 2 --warmup 1 > gpurun_out/s_bench_ref_${NG}gpu.json 2>&1
python -c "
import json,sys
for l in open('gpurun_out/s_bench_${NG}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','clocks')}); print('e2e', d['e2e']['value']); print({k:(v['rollout_ms'],v['update_ms'],v['train_env_steps_per_s']) for k,v in d['training'].items()})"
grep -c impl gpurun_out/s_bench_ref_${NG}gpu.json
