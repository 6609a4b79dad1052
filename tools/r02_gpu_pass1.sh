#!/bin/bash
# round-2 GPU pass 1: tests, headline bench, the two new configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02_p1_pytest.log
echo "pytest rc=$?" >> gpurun_out/r02_p1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_p1_bench.json 2> gpurun_out/r02_p1_bench.err
timeout 600 python bench.py --config yolo26n-seg --steps 10 --warmup 3 > gpurun_out/r02_p1_yolo.json 2> gpurun_out/r02_p1_yolo.err
timeout 300 python bench.py --config tts-decoder --steps 20 --warmup 3 > gpurun_out/r02_p1_tts.json 2> gpurun_out/r02_p1_tts.err
tail -5 gpurun_out/r02_p1_pytest.log; tail -c 600 gpurun_out/r02_p1_bench.json; tail -3 gpurun_out/r02_p1_bench.err; tail -c 400 gpurun_out/r02_p1_yolo.json; tail -3 gpurun_out/r02_p1_yolo.err; tail -c 400 gpurun_out/r02_p1_tts.json; tail -3 gpurun_out/r02_p1_tts.err
