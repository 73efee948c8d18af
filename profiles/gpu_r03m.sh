#!/bin/bash
# round 2, session 4, call M: final state -- full GPU suite, smoke, bench (N=1, default flags) with breakdown, batch-4 bench,
# ncu full capture of the forward GEMMs (student AND teacher variants) after the epilogue changes, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=r03m
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -rfE 2>&1 | tail -20 ) > gpurun_out/m3_pytest.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/m3_smoke.log 2>&1
( time timeout 600 python bench.py --breakdown-file gpurun_out/step_breakdown_$TAG.txt ) > gpurun_out/m3_bench.log 2>&1
( timeout 300 python bench.py --per-gpu-batch 4 --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/step_breakdown_${TAG}_b4.txt ) > gpurun_out/m3_bench_b4.log 2>&1
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step"
cap() { timeout 500 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1; }
cap gemm_fwd 'pw_gemm_kernel<\(int\)256, \(int\)1|pw_gemm_kernel<\(int\)256, \(int\)2' 8 8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/launches_run_$TAG.log 2>&1
grep -n "passed\|failed" gpurun_out/m3_pytest.log; tail -3 gpurun_out/m3_smoke.log; tail -1 gpurun_out/m3_bench.log | cut -c1-700
echo "B=4: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/m3_bench_b4.log | head -2 | tr '\n' ' ')"
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
