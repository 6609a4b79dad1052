#!/bin/bash
# round-2 full pass: every GPU test, the headline bench, the two other configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest.log
grep -E "passed|failed" gpurun_out/r02_pytest.log | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 600 python bench.py --config yolo26n-seg --steps 10 --warmup 3 > gpurun_out/r02_yolo.json 2> gpurun_out/r02_yolo.err
timeout 300 python bench.py --config tts-decoder --steps 20 --warmup 3 > gpurun_out/r02_tts.json 2> gpurun_out/r02_tts.err
python - <<'PY'
import json
for f in ("r02_bench.json","r02_yolo.json","r02_tts.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d[k] for k in ("value","unit","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), "roof", d["roofline"]["achieved"], d["roofline"]["frac"], d.get("exact_mode"), d.get("parity",{}) and d["parity"].get("ids_agreement"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r02_bench.err gpurun_out/r02_yolo.err gpurun_out/r02_tts.err
