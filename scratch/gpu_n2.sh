#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2.log
tail -3 gpurun_out/bench_n2.log | cut -c1-1200

