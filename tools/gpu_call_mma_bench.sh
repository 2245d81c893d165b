#!/bin/bash
# one gpurun call: the tcgen05 issue-rate probe, one process per case (a wrong descriptor may fault)
mkdir -p gpurun_out
n=$(tools/_bin/tc_mma_bench 64 -2 | tail -1)
first=${1:-0}
: > gpurun_out/mma_bench.txt
for i in $(seq $first $((n - 1))); do
  timeout 30 tools/_bin/tc_mma_bench 64 $i >> gpurun_out/mma_bench.txt 2>&1 || echo "case $i: rc=$?" >> gpurun_out/mma_bench.txt
done
cat gpurun_out/mma_bench.txt
