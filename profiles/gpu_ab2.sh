#!/bin/bash
# usage: gpu_ab2.sh "ENV=.. ENV2=.." var ... (pairs: envspec var) -- per-kernel breakdown on the SAME box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
i=0
while [ $# -ge 2 ]; do
  e="$1"; v="$2"; shift 2; i=$((i+1))
  ( env $e FQSS_LIB_PATH=$PWD/fqss_b200/_lib/var/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_ab2_$i.txt ) > gpurun_out/ab2_$i.log 2>&1
  echo "== [$i] $v $e: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab2_$i.log | head -2 | tr '\n' ' ') $(grep 'kernel time sum' gpurun_out/bd_ab2_$i.txt | grep -o 'sum [0-9.]* ms')"
  grep "gln2_dw\|gln1_bwd\|gln2_sums\|tail_bwd\|tcn_dw_fwd\|hidden_fq" gpurun_out/bd_ab2_$i.txt | awk '{printf "   %-22s %s us\n", $1, $4}'
done
