#!/bin/bash
# round 2, call B: persistent-row P1 / P2D kernels -- parity tests + A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_fused.py tests/test_gpu_model.py -m gpu -q --timeout 600 -rfE 2>&1 | tail -60 ) > gpurun_out/b_pytest.log 2>&1
for v in "FQSS_ROWS_PERSIST=0" "FQSS_FR_VAR=0" "FQSS_FR_VAR=1" "FQSS_FR_VAR=2" "FQSS_FR_VAR=0 FQSS_ROW_RJ=2" "FQSS_FR_VAR=0 FQSS_ROW_RJ=8"; do
  echo "== $v" >> gpurun_out/b_ab.log
  ( env $v timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('ms/step %.3f  e2e %.1f  roofline %s %.1fus frac %.3f' % (d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_us'], d['roofline']['frac']))
        for k in d['kernels']: print('   %-28s %.3f ms  n=%s' % (k['kernel'], k['ms_per_step'], k['launches_per_step']))
" ) >> gpurun_out/b_ab.log 2>&1
done
tail -40 gpurun_out/b_pytest.log; cat gpurun_out/b_ab.log
