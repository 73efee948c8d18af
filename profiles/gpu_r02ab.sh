#!/bin/bash
# round 2, call AB: ncu full capture of the block GEMMs in their round-2 form (expand, res/skip, both dgrads, weight gradient)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=r02ab
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step"
cap() { timeout 500 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1; }
cap gemm_fwd 'pw_gemm_kernel<\(int\)256, \(int\)1|pw_gemm_kernel<\(int\)256, \(int\)2' 8 4
cap gemm_bwd 'pw_gemm_kernel<\(int\)256, \(int\)3|pw_gemm_kernel<\(int\)128, \(int\)4|wgrad_kernel' 8 4
ls -la gpurun_out/*_$TAG.ncu-rep
