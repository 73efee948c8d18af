#!/bin/bash
# usage: gpu_ab.sh "<grep regex for breakdown lines>" label1 "ENV1=.. ENV2=.." label2 "..." ...
mkdir -p gpurun_out
RE=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
while [ $# -gt 0 ]; do
  label=$1; envs=$2; shift 2
  env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/bd_$label.txt > gpurun_out/bench_$label.log 2>&1
  echo "== $label [$envs]: $(tail -1 gpurun_out/bench_$label.log | python -c 'import sys,json; print(json.loads(sys.stdin.read())["ms_per_step"])')"
  grep -E "$RE" gpurun_out/bd_$label.txt
done
