#!/bin/bash
# first GPU pass of a session: tests, bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-step > gpurun_out/launches_run.log 2>&1
python profiles/agg_launches.py gpurun_out/launches.csv 40 > gpurun_out/launches_agg.txt 2>&1
head -30 gpurun_out/launches_agg.txt
