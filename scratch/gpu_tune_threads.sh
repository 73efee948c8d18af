#!/bin/bash
# sweep CTA widths of the row kernels (bench breakdown only)
run() { env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --breakdown-file gpurun_out/bd.txt > gpurun_out/b.log 2>&1; echo "== $@: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/b.log | head -1)"; grep "tcn_dw\|tcn_gln\|tcn_hidden\|tcn_tail" gpurun_out/bd.txt | awk '{printf "   %-22s %7.1f\n", $1, $4}'; }
run FQSS_X=0
run FQSS_THREADS_DW_BWD=64 FQSS_THREADS_GLN2_BWD1=64 FQSS_THREADS_GLN2_BWD2=64 FQSS_THREADS_GLN1_BWD=128 FQSS_THREADS_DW_FWD=128 FQSS_THREADS_HIDDEN_FQ=128 FQSS_THREADS_TAIL_BWD=128 FQSS_THREADS_DW_FWD_FLOAT=64 FQSS_THREADS_HIDDEN_FQ_FLOAT=128
run FQSS_THREADS_DW_BWD=96 FQSS_THREADS_GLN2_BWD1=96 FQSS_THREADS_GLN2_BWD2=96 FQSS_THREADS_GLN1_BWD=192 FQSS_THREADS_DW_FWD=192 FQSS_THREADS_HIDDEN_FQ=192 FQSS_THREADS_TAIL_BWD=192 FQSS_THREADS_DW_FWD_FLOAT=96 FQSS_THREADS_HIDDEN_FQ_FLOAT=192
