#!/bin/bash
# round 2, call Z: gLN2 / FQ4 sums in the dgrad GEMM's epilogue -- fused-engine parity suites, A/B bench with breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | grep -v Warn | tail -30 ) > gpurun_out/z_pytest.log 2>&1
tail -6 gpurun_out/z_pytest.log
for v in 1 0; do
  ( FQSS_SUMS_IN_DGRAD=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_z_s$v.txt ) > gpurun_out/z_bench_s$v.log 2>&1
  echo "sums_in_dgrad=$v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/z_bench_s$v.log | head -2 | tr '\n' ' ')"
  grep "gemm_dgrad\|gln2_sums\|gln2_dw" gpurun_out/bd_z_s$v.txt
done
( FQSS_SUMS_IN_DGRAD=1 timeout 300 python bench.py --per-gpu-batch 4 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline ) > gpurun_out/z_bench_b4.log 2>&1
echo "B=4: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/z_bench_b4.log | head -2 | tr '\n' ' ')"
