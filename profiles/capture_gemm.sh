#!/bin/bash
# ncu --set full on the tcgen05 1x1-conv GEMM launches of one steady-state step (student fwd, teacher, bwd).
TAG=${1:-r01}; B=${2:-32}
COMMON="--set full --clock-control none --import-source on --profile-from-start off"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --profile-step"
ncu $COMMON -k regex:pw_gemm_kernel --launch-skip 10 --launch-count 2 -f -o gpurun_out/gemm_fwd_$TAG $BENCH > gpurun_out/gemm_fwd_$TAG.log 2>&1
ncu $COMMON -k regex:pw_gemm_kernel --launch-skip 61 --launch-count 2 -f -o gpurun_out/gemm_teacher_$TAG $BENCH > gpurun_out/gemm_teacher_$TAG.log 2>&1
ncu $COMMON -k regex:pw_gemm_kernel --launch-skip 110 --launch-count 2 -f -o gpurun_out/gemm_bwd_$TAG $BENCH > gpurun_out/gemm_bwd_$TAG.log 2>&1
