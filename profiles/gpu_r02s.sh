#!/bin/bash
# round 2, call S: native LSTM recurrence -- unit test, DPTNet parity, DPTNet bench line with / without it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_dptnet.py -m gpu -q --timeout 300 2>&1 | grep -v Warn | tail -40 ) > gpurun_out/s_new.log 2>&1
tail -40 gpurun_out/s_new.log
( timeout 600 python bench.py --workload dptnet --steps 5 --warmup 3 ) > gpurun_out/s_bench_dptnet.log 2>&1
grep "^{" gpurun_out/s_bench_dptnet.log | cut -c1-2600
( FQSS_NATIVE_LSTM=0 timeout 600 python bench.py --workload dptnet --steps 5 --warmup 3 ) > gpurun_out/s_bench_dptnet_torch.log 2>&1
grep "^{" gpurun_out/s_bench_dptnet_torch.log | cut -c1-2600
