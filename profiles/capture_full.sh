#!/bin/bash
# ncu --set full captures of the hot kernels of one steady-state QAT step (run under gpurun, 1 GPU).
# usage: profiles/capture_full.sh <tag> [per-gpu-batch]
TAG=${1:-r01}; B=${2:-32}
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --profile-step"
# forward kernels of TCN block 2 (skip blocks 0-1)
ncu $COMMON -k 'regex:pw_gemm_kernel<256, [12]>|tcn_dw_fwd|tcn_hidden_fq' --launch-skip 8 --launch-count 4 \
    -f -o gpurun_out/full_fwd_$TAG $BENCH > gpurun_out/full_fwd_$TAG.log 2>&1
# backward kernels of one middle block
ncu $COMMON -k 'regex:tcn_.*bwd_kernel|pw_gemm_kernel<128, 4>|pw_gemm_kernel<256, 3>|wgrad_kernel' --launch-skip 18 --launch-count 9 \
    -f -o gpurun_out/full_bwd_$TAG $BENCH > gpurun_out/full_bwd_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
