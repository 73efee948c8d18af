#!/bin/bash
# round 2, session 4, call J: final-state evidence -- full GPU suite, smoke, bench (N=1) with breakdown, ncu full captures of
# the row kernels (two accumulator copies per sample, 8 quads in flight in the forward kernels), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=r03j
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rfE 2>&1 | tail -20 ) > gpurun_out/j3_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/j3_smoke.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 --breakdown-file gpurun_out/step_breakdown_$TAG.txt ) > gpurun_out/j3_bench.log 2>&1
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step"
cap() {  # name, regex, skip, count
  timeout 300 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1
}
cap rows_bwd 'tcn_tail_bwd_kernel|tcn_gln2_sums_lean_kernel|tcn_gln2_dw_bwd_lean_kernel|tcn_gln1_bwd_kernel' 8 4
cap rows_fwd 'tcn_dw_fwd_kernel<\(bool\)1|tcn_hidden_fq_kernel<\(bool\)1' 4 2
cap rows_fwd_float 'tcn_dw_fwd_kernel<\(bool\)0' 4 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/launches_run_$TAG.log 2>&1
tail -4 gpurun_out/j3_pytest.log; tail -6 gpurun_out/j3_smoke.log; tail -1 gpurun_out/j3_bench.log | cut -c1-900
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
