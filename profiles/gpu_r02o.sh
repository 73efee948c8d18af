#!/bin/bash
# round 2, call O: DPTNetQ on the quantiser kernels vs the reference's golden vectors
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dptnet.py tests/test_gpu_sepformer.py -m gpu -q --timeout 300 2>&1 | tail -60 ) > gpurun_out/o_new.log 2>&1
tail -60 gpurun_out/o_new.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/parity_r02.json")) if __import__("os").path.exists("gpurun_out/parity_r02.json") else {}
print({k: v for k, v in d.items() if k.startswith(("dptnet", "sepformer"))})
PY
