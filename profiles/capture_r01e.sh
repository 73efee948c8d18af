#!/bin/bash
# Final round-1 ncu captures (after: FQ1 codes from the expand GEMM's epilogue, 128-thread depthwise kernels, pitch-
# specialised GEMM epilogues).  Same recipe as capture_r01d.sh; the teacher-GEMM and filterbank captures of r01d still
# describe the current kernels and are not repeated.  usage: profiles/capture_r01e.sh [tag] [per-gpu-batch]
TAG=${1:-r01e}; B=${2:-32}
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --profile-step"
cap() {  # name, regex, skip, count
  timeout 300 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1
}
cap rows_fwd     'tcn_dw_fwd_kernel|tcn_hidden_fq_kernel' 4 2
cap rows_teacher 'tcn_dw_fwd_kernel' 26 1
cap rows_bwd     'tcn_tail_bwd_kernel|tcn_gln2_sums_codes_kernel|tcn_gln2_dw_bwd_kernel|tcn_gln1_bwd_kernel' 8 4
cap gemm_fwd     'pw_gemm_kernel' 6 2
cap gemm_bwd     'pw_gemm_kernel|wgrad_kernel' 106 4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/launches_run_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
