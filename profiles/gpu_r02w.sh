#!/bin/bash
# round 2, call W: block-backward tail on the side stream -- parity suites that drive the fused engine, A/B bench at B = 32 and B = 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | grep -v Warn | tail -12 ) > gpurun_out/w_pytest.log 2>&1
tail -5 gpurun_out/w_pytest.log
for v in 1 0; do
  for b in 32 4; do
    ( FQSS_BWD_TAIL_SIDE=$v timeout 300 python bench.py --per-gpu-batch $b --steps 20 --warmup 5 --no-cpu-baseline --no-roofline ) > gpurun_out/w_bench_t${v}_b$b.log 2>&1
    echo "tail_side=$v B=$b: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/w_bench_t${v}_b$b.log | head -2 | tr '\n' ' ')"
  done
done
