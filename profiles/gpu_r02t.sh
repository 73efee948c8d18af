#!/bin/bash
# round 2, call T: fused attention core -- unit tests, DPTNet / Sepformer parity, both bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_lstm.py tests/test_gpu_dptnet.py tests/test_gpu_sepformer.py -m gpu -q --timeout 300 2>&1 | grep -v Warn | tail -40 ) > gpurun_out/t_new.log 2>&1
tail -30 gpurun_out/t_new.log
for w in dptnet sepformer; do
  ( timeout 600 python bench.py --workload $w --steps 5 --warmup 3 ) > gpurun_out/t_bench_$w.log 2>&1
  grep "^{" gpurun_out/t_bench_$w.log | cut -c1-2400
  ( FQSS_NATIVE_ATTENTION=0 timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-roofline ) > gpurun_out/t_bench_${w}_torchattn.log 2>&1
  grep "^{" gpurun_out/t_bench_${w}_torchattn.log | cut -c1-300
done
