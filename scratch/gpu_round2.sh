#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown-file gpurun_out/breakdown.txt > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1500
cat gpurun_out/breakdown.txt
timeout 1200 bash profiles/capture_full.sh r01a 32
