#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_resident.py tests/test_gpu_parity.py tests/test_gpu_zz_model_forms.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02_resident_pytest.log
tail -3 gpurun_out/r02_resident_pytest.log
python tools/yolo_op_profile.py 2>&1 | tail -48 | head -30
timeout 600 python bench.py --config yolo26n-seg --steps 10 --warmup 3 > gpurun_out/r02_yolo_fold.json 2> gpurun_out/r02_yolo_fold.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_yolo_fold.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","unit","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"])
PY
tail -3 gpurun_out/r02_yolo_fold.err
