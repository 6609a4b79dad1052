#!/bin/bash
# quick A/B of the int8 GEMM variants on a short stack + the in-kernel role timelines (rebuilds gemm_i8_tc.cu with -DLELE_B200_GEMM_TIMELINE)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-tl}
{
timeout 300 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q -k "one_pass or bit_exact or lanes" 2>&1 | tail -3
echo "== product build"; QS_LAYERS=8 timeout 300 python tools/quick_step.py
for v in ${QS_VARIANTS:-LELE_B200_FFN_RESB=0 LELE_B200_FFN_FUSED=0}; do echo "== $v"; env $v QS_LAYERS=8 timeout 300 python tools/quick_step.py; done
touch lele_b200/csrc/gemm_i8_tc.cu
LELE_B200_NVCC_DEFS=-DLELE_B200_GEMM_TIMELINE python lele_b200/build.py > /dev/null
echo "== timeline build (eager, 2 layers)"
LELE_B200_GEMM_DBG=1 LELE_B200_GRAPH=0 QS_LAYERS=2 timeout 300 python tools/quick_step.py 2>&1 | grep -E "FQDBG|GEMMDBG|QS" | tail -60
} > gpurun_out/${tag}.log 2>&1
tail -80 gpurun_out/${tag}.log
