#!/bin/bash
# the headline bench + ncu launch list (time + DRAM bytes) + the switch tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 600 python -m pytest tests/test_gpu_sensevoice.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^(?!prep_).*" -c 760 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exact-mode > gpurun_out/launches_$tag.log 2>&1
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"], d["parity"]["ids_agreement"], d["clocks"])
print(d["kernel_breakdown_ms"])
PY
