#!/bin/bash
# round 2, session 4, call K: compute-sanitizer memcheck over the WHOLE GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 2700 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q --timeout 2400 -p no:cacheprovider > gpurun_out/k3_memcheck_all.log 2>&1 ); echo "rc=$?"
grep -n "========= [A-Z]" gpurun_out/k3_memcheck_all.log | cut -c1-200 | head -30
grep -c "Invalid __\|out of bounds\|misaligned" gpurun_out/k3_memcheck_all.log
tail -4 gpurun_out/k3_memcheck_all.log
