#!/bin/bash
# round 2c: SenseVoice GPU tests, A/B of the switches named in $QS_VARIANTS on a short stack, then the headline bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02c}
{
timeout 1200 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q ${QS_PYTEST_K:+-k "$QS_PYTEST_K"} 2>&1 | tail -15
echo "== default"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
for v in ${QS_VARIANTS}; do echo "== $v"; env ${v//,/ } QS_LAYERS=8 timeout 300 python tools/quick_step.py; done
} > gpurun_out/${tag}.log 2>&1
grep -E "passed|failed|rror|QS|layer_norm=|==" gpurun_out/${tag}.log | cut -c1-700
if [ -z "$QS_NO_BENCH" ]; then
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
fi
