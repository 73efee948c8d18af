#!/bin/bash
# ncu --set full of selected kernels of one steady-state step. usage: gpu_ncu_one.sh <tag> <regex> <skip> <count>
TAG=$1; RE=$2; SKIP=${3:-10}; CNT=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:$RE" --launch-skip $SKIP --launch-count $CNT -f -o gpurun_out/one_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step > gpurun_out/one_$TAG.log 2>&1
ls -la gpurun_out/one_$TAG.ncu-rep
