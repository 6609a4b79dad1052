#!/bin/bash
# role timelines of conv_tc_pixel_kernel (clock64 wait counters printed by CTA 0): rebuild conv_tc.cu with -DLELE_B200_CONV_TIMELINE
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
touch lele_b200/csrc/conv_tc.cu
LELE_B200_NVCC_DEFS=-DLELE_B200_CONV_TIMELINE python lele_b200/build.py > /dev/null
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from lele_b200 import Context, kernels as K
ctx = Context(0)
rng = np.random.default_rng(0)
for (nb, ic, hw, oc, k, st, act) in [(32,48,160,64,1,1,0),(32,48,160,64,1,1,2),(32,48,160,16,1,1,0),(32,64,160,64,3,1,0),(32,3,640,16,3,2,0)]:
    x = ctx.to_device(rng.standard_normal((nb, ic, hw, hw)).astype(np.float32))
    w = ctx.to_device((rng.standard_normal((oc, ic, k, k)) / np.sqrt(ic*k*k)).astype(np.float32)); b = ctx.to_device(rng.standard_normal(oc).astype(np.float32))
    p = k // 2
    print(f"--- nb{nb} ic{ic} hw{hw} oc{oc} k{k} s{st} act{act}", flush=True)
    K.conv2d(x, w, b, (1,1), 1, (p,p,p,p), (st,st), act, ctx=ctx); ctx.sync()
PY
touch lele_b200/csrc/conv_tc.cu
