#!/bin/bash
# round 2, call G: instruction-lean per-row fused gLN2+depthwise backward -- parity tests, A/B timing, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 -rfE 2>&1 | tail -25 ) > gpurun_out/g_pytest.log 2>&1
rm -f gpurun_out/g_ab.log
for v in "FQSS_FL_VAR=0" "FQSS_LEAN_P2D=0"; do
  echo "== $v" >> gpurun_out/g_ab.log
  ( env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('ms/step %.3f  e2e %.1f  roofline %s %.1fus frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac']))
        for k in d['kernels'][:12]: print('   %-28s %.3f ms  n=%s' % (k['kernel'], k['ms_per_step'], k['launches_per_step']))
" ) >> gpurun_out/g_ab.log 2>&1
done
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch 32 --no-cpu-baseline --no-roofline --profile-step"
timeout 400 ncu $COMMON -k "regex:tcn_gln2_sums_lean_kernel" --launch-skip 2 --launch-count 1 -f -o gpurun_out/lean_p1 $BENCH > gpurun_out/lean_p1.log 2>&1
tail -12 gpurun_out/g_pytest.log; cat gpurun_out/g_ab.log
