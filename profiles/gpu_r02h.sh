#!/bin/bash
# round 2, call H: full GPU suite with the reassociation yardstick + smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/parity_r02.json
( time timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -rfE --durations=8 2>&1 | tail -60 ) > gpurun_out/h_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/h_smoke.log 2>&1
tail -25 gpurun_out/h_pytest.log | cut -c1-400; tail -6 gpurun_out/h_smoke.log
