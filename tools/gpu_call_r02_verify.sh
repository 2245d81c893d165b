#!/bin/bash
# Short verification of the tree as it stands: all GPU tests, smoke, default bench + reference arm.  Outputs gpurun_out/v_*.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -4 gpurun_out/v_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/v_smoke.log 2>&1; tail -2 gpurun_out/v_smoke.log
timeout 900 python bench.py --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -2 gpurun_out/v_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v_bench_ref.json 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/v_bench.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')})
        print('e2e', {k: v for k, v in d['e2e'].items() if not isinstance(v, str)})
        print('roofline frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'])
        print(d['training'].get('bf16x3'))
PY
