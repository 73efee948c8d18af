#!/bin/bash
# per-kernel breakdown at several per-GPU batch sizes: does a smaller working set (L2-resident producer->consumer
# tensors) shorten the kernels more than proportionally?
mkdir -p gpurun_out
for B in 2 4 8 16; do
  timeout 300 python bench.py --steps 10 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --breakdown-file gpurun_out/breakdown_B$B.txt > gpurun_out/bench_B$B.log 2>&1
  tail -1 gpurun_out/bench_B$B.log | cut -c1-200
done
