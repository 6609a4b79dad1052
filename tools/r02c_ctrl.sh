#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q 2>&1 | tail -6
echo "== control warps high (default)"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
QS_LAYERS=8 timeout 300 python tools/quick_step.py | head -1
touch lele_b200/csrc/gemm_i8_tc.cu lele_b200/csrc/attn_tc.cu
LELE_B200_NVCC_DEFS=-DLELE_B200_CTRL_WARPS_LOW python lele_b200/build.py > /dev/null
echo "== control warps low (round 1 placement)"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
QS_LAYERS=8 timeout 300 python tools/quick_step.py | head -1
} > gpurun_out/r02c_ctrl.log 2>&1
grep -E "passed|failed|rror|QS|layer_norm=|==" gpurun_out/r02c_ctrl.log | cut -c1-700
