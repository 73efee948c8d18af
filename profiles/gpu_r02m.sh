#!/bin/bash
# round 2, call M: mask-head fusion -- new kernel tests, whole suite, bench with breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout 300 -x -k mask_head 2>&1 | tail -30 ) > gpurun_out/m_new.log 2>&1
tail -30 gpurun_out/m_new.log
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 ) > gpurun_out/m_pytest.log 2>&1
tail -8 gpurun_out/m_pytest.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/bd_m.txt ) > gpurun_out/m_bench.log 2>&1
grep "^{" gpurun_out/m_bench.log | cut -c1-300; sed -n 1,8p gpurun_out/bd_m.txt; grep -n "mask_head\|pw_bwd\|pw_fwd\|gemm_store\|gemm_relu" gpurun_out/bd_m.txt
