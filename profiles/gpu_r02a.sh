#!/bin/bash
# round 2, call A: full GPU test suite (new kernel-level parity tests), smoke, baseline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
free -g >> gpurun_out/a_smi.txt; nproc >> gpurun_out/a_smi.txt
( time timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -rfE --durations=15 2>&1 | tail -250 ) > gpurun_out/a_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/a_bench.log 2>&1
tail -30 gpurun_out/a_pytest.log; tail -8 gpurun_out/a_smoke.log; tail -2 gpurun_out/a_bench.log | cut -c1-600
