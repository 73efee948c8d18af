#!/bin/bash
# round 2, call U: final evidence -- bench with breakdown, ncu full capture of the new edge kernels, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=r02u
( timeout 600 python bench.py --steps 10 --warmup 3 --breakdown-file gpurun_out/step_breakdown_$TAG.txt ) > gpurun_out/u_bench.log 2>&1
grep "^{" gpurun_out/u_bench.log | cut -c1-900
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step"
cap() {  # name, regex, skip, count
  timeout 400 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1
}
cap edge 'pw_gemm_kernel<128, 0|sub_fq_codes_kernel|frames_split_kernel|frames_encode_kernel|ola_fwd_kernel|dec_prep_kernel|mask_head_bwd_kernel' 0 14
cap maskhead 'pw_gemm_kernel<256, 5' 0 2
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/launches_run_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
