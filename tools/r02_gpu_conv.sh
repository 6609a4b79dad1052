#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/conv_microbench.py 2>&1 | tail -14
timeout 400 python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py tests/test_gpu_zz_model_forms.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_resident_pytest.log
tail -3 gpurun_out/r02_resident_pytest.log
timeout 200 python tools/yolo_op_profile.py 2>&1 | grep -E "step ms|^conv"
