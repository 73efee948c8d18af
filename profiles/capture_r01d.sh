#!/bin/bash
# ncu captures of one steady-state QAT step (bench.py --profile-step brackets it with cudaProfilerStart/Stop), 1 GPU.
# usage: profiles/capture_r01d.sh <tag> [per-gpu-batch]     (run under gpurun; reports land in gpurun_out/)
TAG=${1:-r01d}; B=${2:-32}
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
BENCH="python bench.py --steps 1 --warmup 3 --per-gpu-batch $B --no-cpu-baseline --no-roofline --profile-step"
cap() {  # name, regex, skip, count
  timeout 600 ncu $COMMON -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o gpurun_out/$1_$TAG $BENCH > gpurun_out/$1_$TAG.log 2>&1
}
# student forward row kernels of TCN block 2; the teacher's depthwise kernel (launches 24.. of that name)
cap rows_fwd     'tcn_dw_fwd_kernel|tcn_hidden_fq_kernel' 4 2
cap rows_teacher 'tcn_dw_fwd_kernel' 26 1
# backward row kernels of one middle block: tail, gLN2 sums (codes), fused gLN2+depthwise, gLN1
cap rows_bwd     'tcn_tail_bwd_kernel|tcn_gln2_sums_codes_kernel|tcn_gln2_dw_bwd_kernel|tcn_gln1_bwd_kernel' 8 4
# tcgen05 GEMMs: student fwd (expand, res|skip), teacher (split-bf16), backward (dgrad / wgrad)
cap gemm_fwd     'pw_gemm_kernel' 6 2
cap gemm_teacher 'pw_gemm_kernel' 55 2
cap gemm_bwd     'pw_gemm_kernel|wgrad_kernel' 106 4
# filterbank (encoder / decoder / RQB) kernels
cap edge         'edge_wgrad_lanes_kernel|analysis_fwd_kernel|synthesis_fwd_kernel' 0 8
# launch list of the whole step (durations only; cold-cache + serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  $BENCH > gpurun_out/launches_run_$TAG.log 2>&1
ls -la gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
