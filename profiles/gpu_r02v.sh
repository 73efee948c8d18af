#!/bin/bash
# round 2, call V: compute-sanitizer (memcheck + racecheck) over the kernels added in this session
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="tests/test_gpu_edge_tc.py tests/test_gpu_lstm.py tests/test_gpu_attention.py"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -q --timeout 1200 -x 2>&1 | tail -40 ) > gpurun_out/v_memcheck.log 2>&1
echo "memcheck:"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/v_memcheck.log; tail -6 gpurun_out/v_memcheck.log
( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_attention.py -m gpu -q --timeout 1200 -x -k "not mha_q" 2>&1 | tail -40 ) > gpurun_out/v_racecheck.log 2>&1
echo "racecheck:"; grep -c "hazard" gpurun_out/v_racecheck.log; tail -6 gpurun_out/v_racecheck.log
