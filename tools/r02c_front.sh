#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frontend or cmvn" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q 2>&1 | tail -8
echo "== default"; QS_LAYERS=4 timeout 300 python tools/quick_step.py
for v in ${QS_VARIANTS}; do echo "== $v"; env ${v//,/ } QS_LAYERS=4 timeout 300 python tools/quick_step.py; done
} > gpurun_out/r02c_front.log 2>&1
grep -E "passed|failed|rror|QS|layer_norm=|==|assert|Mismatch|Max " gpurun_out/r02c_front.log | cut -c1-400
