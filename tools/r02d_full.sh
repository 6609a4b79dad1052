#!/bin/bash
# round-2 final pass: every GPU test, the three bench configs + reference arm, ncu launch list (time + DRAM bytes) of one forward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02d}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${tag}_pytest.log
grep -E "passed|failed|rror" gpurun_out/${tag}_pytest.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 600 python bench.py --config yolo26n-seg --steps 10 --warmup 3 > gpurun_out/${tag}_yolo.json 2> gpurun_out/${tag}_yolo.err
timeout 300 python bench.py --config tts-decoder --steps 20 --warmup 3 > gpurun_out/${tag}_tts.json 2> gpurun_out/${tag}_tts.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^(?!prep_).*" -c 760 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exact-mode > gpurun_out/launches_$tag.log 2>&1
python - <<PY
import json
for f in ("${tag}_bench.json","${tag}_bench_ref.json","${tag}_yolo.json","${tag}_tts.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","unit","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "roof", d.get("roofline",{}).get("achieved"), d.get("roofline",{}).get("frac"), d.get("exact_mode"), (d.get("parity") or {}).get("ids_agreement"), d.get("clocks"))
        if "kernel_breakdown_ms" in d: print(d["kernel_breakdown_ms"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/${tag}_bench.err gpurun_out/${tag}_yolo.err gpurun_out/${tag}_tts.err; wc -l gpurun_out/launches_$tag.csv
