#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/bd_$label.txt > gpurun_out/bench_$label.log 2>&1
  echo "== $label: $(tail -1 gpurun_out/bench_$label.log | python -c 'import sys,json; print(json.loads(sys.stdin.read())["ms_per_step"])')"
  grep -E "tcn_dw_bwd|tcn_gln1_bwd|tcn_gln2_bwd<2>|tcn_hidden_fq " gpurun_out/bd_$label.txt
}
run a FQSS_NQ_DW=1 FQSS_NQ_P2=1 FQSS_NQ_Q=1
run b FQSS_NQ_DW=2 FQSS_NQ_P2=2 FQSS_NQ_Q=2
run c FQSS_NQ_DW=4 FQSS_NQ_P2=4 FQSS_NQ_Q=4
run d FQSS_NQ_DW=4 FQSS_NQ_P2=2 FQSS_NQ_Q=12
run e FQSS_NQ_DW=4 FQSS_NQ_P2=2 FQSS_NQ_Q=14
