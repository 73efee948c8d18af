#!/bin/bash
# round 2, call C: ncu --set full of the persistent-row kernels (two register variants)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch 32 --no-cpu-baseline --no-roofline --profile-step"
for v in 0 2; do
  FQSS_FR_VAR=$v timeout 400 ncu $COMMON -k "regex:tcn_gln2_sums_rows_kernel|tcn_gln2_dw_bwd_rows_kernel" --launch-skip 4 --launch-count 2 -f -o gpurun_out/rows_var$v $BENCH > gpurun_out/rows_var$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
