#!/bin/bash
# usage: gpu_ab_lean.sh "<grep regex>" label1 "ENV.." label2 "ENV.." ...   (no pytest; eager breakdown only)
mkdir -p gpurun_out
RE=$1; shift
while [ $# -gt 0 ]; do
  label=$1; envs=$2; shift 2
  env $envs timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --cuda-graph 0 --breakdown-file gpurun_out/bd_$label.txt > gpurun_out/bench_$label.log 2>&1
  echo "== $label [$envs]: $(tail -1 gpurun_out/bench_$label.log | python -c 'import sys,json; print(json.loads(sys.stdin.read())["ms_per_step"])') $(sed -n 2p gpurun_out/bd_$label.txt | cut -c50-)"
  grep -E "$RE" gpurun_out/bd_$label.txt
done
