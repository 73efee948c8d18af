#!/bin/bash
# usage: gpu_ab.sh var1 var2 ...   -- per-kernel breakdown of every variant library on the SAME box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in "$@"; do
  ( FQSS_LIB_PATH=$PWD/fqss_b200/_lib/var/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_ab_$v.txt ) > gpurun_out/ab_$v.log 2>&1
  echo "== $v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$v.log | head -2 | tr '\n' ' ') $(grep 'kernel time sum' gpurun_out/bd_ab_$v.txt | grep -o 'sum [0-9.]* ms')"
  grep "gln2_dw\|gln1_bwd\|gln2_sums\|tail_bwd\|tcn_dw_fwd \|hidden_fq" gpurun_out/bd_ab_$v.txt | awk '{printf "   %-22s %s us\n", $1, $4}'
done
