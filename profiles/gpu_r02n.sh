#!/bin/bash
# round 2, call N: filterbank edges on the tcgen05 GEMM -- new tests, whole suite, bench with breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_edge_tc.py -m gpu -q --timeout 300 2>&1 | tail -60 ) > gpurun_out/n_new.log 2>&1
tail -60 gpurun_out/n_new.log
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 ) > gpurun_out/n_pytest.log 2>&1
tail -12 gpurun_out/n_pytest.log
( timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/bd_n.txt ) > gpurun_out/n_bench.log 2>&1
grep "^{" gpurun_out/n_bench.log | cut -c1-400; tail -5 gpurun_out/n_bench.log | cut -c1-300; sed -n 1,12p gpurun_out/bd_n.txt
grep -n "frames\|ola\|sub_fq\|edge\|sconv\|tconv\|gemm_store\|wgrad\|mask_head\|gemm_relu\|pw_bwd\|pw_fwd\|dec_wgrad" gpurun_out/bd_n.txt
