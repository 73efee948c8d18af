#!/bin/bash
# round 2, call I (2 GPUs): data-parallel bench line with dp_check + strong sub-record, clean teardown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/i_bench2.log 2>&1
echo "exit code $?" >> gpurun_out/i_bench2.log
grep -E "^\{|exit code|real|Error|error" gpurun_out/i_bench2.log | cut -c1-2500
