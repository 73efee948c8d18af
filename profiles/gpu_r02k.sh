#!/bin/bash
# round 2, call K: ConvTasNetMusicQ CUDA path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_music.py -m gpu -q --timeout 600 -rfE 2>&1 | tail -60 ) > gpurun_out/k_pytest.log 2>&1
tail -40 gpurun_out/k_pytest.log | cut -c1-600
