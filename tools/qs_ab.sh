#!/bin/bash
# quick A/B on a short stack: tests named by $QS_TESTS, then tools/quick_step.py under each env setting in $QS_VARIANTS
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-ab}
{
timeout 600 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q -k "${QS_TESTS:-one_pass or bit_exact or lanes}" 2>&1 | tail -5
echo "== default"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
for v in ${QS_VARIANTS}; do echo "== $v"; env $v QS_LAYERS=8 timeout 300 python tools/quick_step.py; done
} > gpurun_out/${tag}.log 2>&1
grep -E "passed|failed|rror|QS|layer_norm=|==" gpurun_out/${tag}.log | cut -c1-600
