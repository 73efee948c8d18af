#!/bin/bash
# round 2, call Q: per-kernel breakdown at the per-GPU share of BASELINE configs[1] on 8 GPUs (batch 4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 300 python bench.py --per-gpu-batch 4 --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_q_b4.txt ) > gpurun_out/q_bench_b4.log 2>&1
grep "^{" gpurun_out/q_bench_b4.log | cut -c1-300; sed -n 1,60p gpurun_out/bd_q_b4.txt
