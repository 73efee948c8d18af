#!/bin/bash
# usage: gpu_ab3.sh grep-pattern var...
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
pat="$1"; shift
i=0
for v in "$@"; do
  i=$((i+1))
  ( FQSS_LIB_PATH=$PWD/fqss_b200/_lib/var/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown-file gpurun_out/bd_ab3_$i.txt ) > gpurun_out/ab3_$i.log 2>&1
  echo "== [$i] $v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab3_$i.log | head -2 | tr '\n' ' ') $(grep 'kernel time sum' gpurun_out/bd_ab3_$i.txt | grep -o 'sum [0-9.]* ms')"
  grep -E "$pat" gpurun_out/bd_ab3_$i.txt | awk '{printf "   %-22s %s us\n", $1, $4}'
done
