#!/bin/bash
# round 2, call D: hand-tuned persistent-row kernels -- parity tests, A/B timing, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 -rfE 2>&1 | tail -25 ) > gpurun_out/f_pytest.log 2>&1
rm -f gpurun_out/f_ab.log
for v in "FQSS_FR_VAR=0" "FQSS_FR_VAR=1" "FQSS_FR_VAR=2"; do
  echo "== $v" >> gpurun_out/f_ab.log
  ( env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('ms/step %.3f  e2e %.1f  roofline %s %.1fus frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac']))
        for k in d['kernels'][:12]: print('   %-28s %.3f ms  n=%s' % (k['kernel'], k['ms_per_step'], k['launches_per_step']))
" ) >> gpurun_out/f_ab.log 2>&1
done
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch 32 --no-cpu-baseline --no-roofline --profile-step"
for v in 0; do
  FQSS_FR_VAR=$v timeout 400 ncu $COMMON -k "regex:tcn_gln2_sums_rows_kernel|tcn_gln2_dw_bwd_rows_kernel" --launch-skip 4 --launch-count 2 -f -o gpurun_out/rowsF_var$v $BENCH > gpurun_out/rowsF_var$v.log 2>&1
done
tail -12 gpurun_out/f_pytest.log; cat gpurun_out/f_ab.log
