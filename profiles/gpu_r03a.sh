#!/bin/bash
# round 2, session 4, call A: trimmed gLN2 sums kernel + depthwise forward (frame checks out of the loops, SHF+LOP3 table
# addresses, IDP.4A code sums) -- fused-engine parity suites, bench with per-kernel breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 2>&1 | grep -v Warn | tail -30 ) > gpurun_out/a3_pytest.log 2>&1
tail -6 gpurun_out/a3_pytest.log
( timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_a3.txt ) > gpurun_out/a3_bench.log 2>&1
echo "B=32: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/a3_bench.log | head -2 | tr '\n' ' ')"
head -24 gpurun_out/bd_a3.txt
( timeout 300 python bench.py --per-gpu-batch 4 --steps 20 --warmup 5 --no-cpu-baseline --no-roofline ) > gpurun_out/a3_bench_b4.log 2>&1
echo "B=4: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/a3_bench_b4.log | head -2 | tr '\n' ' ')"
