#!/bin/bash
# round 2, call J: per-kernel breakdown of the current step (events inside real steps) + launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/step_breakdown_r02a.txt > gpurun_out/j_bench.log 2>&1
cat gpurun_out/step_breakdown_r02a.txt
