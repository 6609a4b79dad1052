#!/bin/bash
# one GPU visit of round 2b: the GEMM-related GPU tests, then the headline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02b}
timeout 900 python -m pytest tests/test_gpu_sensevoice.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --no-exact-mode > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["parity"]["ids_agreement"])
    print(d["kernel_breakdown_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
