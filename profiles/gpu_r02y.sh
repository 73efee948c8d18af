#!/bin/bash
# round 2, call Y: ncu full capture of the GEMM launches of the tensor-core edge path and the mask head
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=r02y
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step"
timeout 500 ncu $COMMON -k 'regex:pw_gemm_kernel<\(int\)128, \(int\)0|pw_gemm_kernel<\(int\)256, \(int\)0|pw_gemm_kernel<\(int\)256, \(int\)4|pw_gemm_kernel<\(int\)256, \(int\)5' --launch-count 13 -f -o gpurun_out/edge_gemm_$TAG $BENCH > gpurun_out/edge_gemm_$TAG.log 2>&1
tail -3 gpurun_out/edge_gemm_$TAG.log; ls -la gpurun_out/edge_gemm_$TAG.ncu-rep
