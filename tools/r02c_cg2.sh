#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== parity, cta_group::2 forced on every shape"
LELE_B200_GEMM_CG2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantized or linear or integer or prepare" 2>&1 | tail -12
LELE_B200_GEMM_CG2=1 timeout 600 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q -k "per_layer or pcm_to_ids or simt_attention" 2>&1 | tail -12
echo "== sensevoice tests, default"
timeout 900 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q 2>&1 | tail -12
echo "== default"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
for v in ${QS_VARIANTS}; do echo "== $v"; env ${v//,/ } QS_LAYERS=8 timeout 300 python tools/quick_step.py; done
} > gpurun_out/r02c_cg2.log 2>&1
grep -E "passed|failed|rror|QS|layer_norm=|==|assert|timeout|Mismatch" gpurun_out/r02c_cg2.log | cut -c1-600
