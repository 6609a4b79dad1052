#!/bin/bash
# in-kernel role timelines of the int8 GEMMs (rebuilds gemm_i8_tc.cu with -DLELE_B200_GEMM_TIMELINE on the GPU box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02c_tl}
touch lele_b200/csrc/gemm_i8_tc.cu
LELE_B200_NVCC_DEFS=-DLELE_B200_GEMM_TIMELINE python lele_b200/build.py > /dev/null
LELE_B200_GEMM_DBG=1 LELE_B200_GRAPH=0 QS_LAYERS=2 timeout 300 python tools/quick_step.py 2>&1 | grep -E "FQDBG|GEMMDBG|GEMMCTA|QS" > gpurun_out/${tag}.log
grep -c GEMMCTA gpurun_out/${tag}.log
grep -E "FQDBG|GEMMDBG" gpurun_out/${tag}.log | tail -150
