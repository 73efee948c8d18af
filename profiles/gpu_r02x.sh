#!/bin/bash
# round 2, call X: row constants folded into the depthwise kernel for small grids -- fused-engine suites, A/B at B = 4 (and B = 32 unchanged)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py tests/test_gpu_edge_tc.py tests/test_gpu_ops.py -m gpu -q --timeout 600 2>&1 | grep -v Warn | tail -12 ) > gpurun_out/x_pytest.log 2>&1
tail -4 gpurun_out/x_pytest.log
for v in 4096 0; do
  for b in 4 8; do
    ( FQSS_RC_FOLD_ROWS=$v timeout 300 python bench.py --per-gpu-batch $b --steps 20 --warmup 5 --no-cpu-baseline --no-roofline ) > gpurun_out/x_bench_f${v}_b$b.log 2>&1
    echo "fold_rows=$v B=$b: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/x_bench_f${v}_b$b.log | head -2 | tr '\n' ' ')"
  done
done
