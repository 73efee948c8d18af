#!/bin/bash
# round 2, call P: extra bench lines of the sequence models (BASELINE configs[2] / [3])
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for w in dptnet sepformer; do
  ( timeout 600 python bench.py --workload $w --steps 5 --warmup 3 ) > gpurun_out/p_bench_$w.log 2>&1
  grep "^{" gpurun_out/p_bench_$w.log | cut -c1-1200 || true
  grep -v "^{" gpurun_out/p_bench_$w.log | grep -v Warning | tail -5
done
