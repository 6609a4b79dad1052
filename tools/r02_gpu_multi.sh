#!/bin/bash
# N-GPU bench (clip-sharded, library NCCL comm): usage tools/r02_gpu_multi.sh N
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_n$N.json; tail -5 gpurun_out/r02_bench_n$N.err
