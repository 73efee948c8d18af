#!/bin/bash
# A/B in graph mode: label env pairs; prints ms/step (value, e2e)
mkdir -p gpurun_out
while [ $# -gt 0 ]; do
  label=$1; envs=$2; shift 2
  for rep in 1 2; do
    env $envs timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_$label.log 2>&1
    echo "== $label [$envs] rep $rep: $(tail -1 gpurun_out/bench_$label.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_mode"][:20])')"
  done
done
