#!/bin/bash
mkdir -p gpurun_out
for g in 0 1; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline --cuda-graph $g > gpurun_out/bench_g$g.log 2>&1; echo "rc=$?"
  tail -1 gpurun_out/bench_g$g.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"], d["gpu_launches"], d.get("launch_mode"), d["final_loss"])' || tail -5 gpurun_out/bench_g$g.log
done
