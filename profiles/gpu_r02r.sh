#!/bin/bash
# round 2, call R: state check -- smoke(), whole GPU suite, default bench (with cpu baseline), reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r_smoke.log 2>&1; echo "smoke rc=$?"; tail -12 gpurun_out/r_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 ) > gpurun_out/r_pytest.log 2>&1
tail -6 gpurun_out/r_pytest.log
( timeout 600 python bench.py ) > gpurun_out/r_bench.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r_bench.log | cut -c1-2500
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r_bench_ref.log 2>&1; echo "ref rc=$?"
grep "^{" gpurun_out/r_bench_ref.log | cut -c1-600
