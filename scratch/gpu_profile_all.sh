#!/bin/bash
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 1500 bash profiles/capture_full.sh $TAG 32
timeout 1500 bash profiles/capture_gemm.sh $TAG 32
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step > gpurun_out/launches_run.log 2>&1
python profiles/agg_launches.py gpurun_out/launches_$TAG.csv 60 > gpurun_out/launches_agg_$TAG.txt 2>&1
