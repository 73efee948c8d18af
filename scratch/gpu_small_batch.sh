#!/bin/bash
mkdir -p gpurun_out
for B in 4 8 16; do
  for g in 1 0; do
    timeout 120 python bench.py --steps 20 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --cuda-graph $g > gpurun_out/bench_sb.log 2>&1
    echo "B=$B graph=$g: $(tail -1 gpurun_out/bench_sb.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],2), "ms/step", round(d["value"],1), "audio-s/s; e2e", round(d["e2e"]["value"],1))')"
  done
done
