#!/bin/bash
# ncu --set full on the tcgen05 GEMM launches of one steady-state step (student fwd, teacher fwd, student bwd).
# Launch order of pw_gemm_kernel inside the profiled step: 48 student fwd (expand, res|skip alternating),
# 1 + 48 + 1 teacher (split-bf16), then per backward block: dgrad(bf16), [wgrad], dgrad(add), [wgrad].
TAG=${1:-r01}; B=${2:-32}
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --profile-step"
ncu $COMMON -k 'regex:pw_gemm_kernel' --launch-skip 6 --launch-count 2 -f -o gpurun_out/gemm_fwd_$TAG $BENCH > gpurun_out/gemm_fwd_$TAG.log 2>&1
ncu $COMMON -k 'regex:pw_gemm_kernel' --launch-skip 55 --launch-count 2 -f -o gpurun_out/gemm_teacher_$TAG $BENCH > gpurun_out/gemm_teacher_$TAG.log 2>&1
ncu $COMMON -k 'regex:pw_gemm_kernel|wgrad_kernel' --launch-skip 106 --launch-count 4 -f -o gpurun_out/gemm_bwd_$TAG $BENCH > gpurun_out/gemm_bwd_$TAG.log 2>&1
ls -la gpurun_out/gemm_*_$TAG.ncu-rep
