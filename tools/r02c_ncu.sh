#!/bin/bash
# one ncu --set full capture of the layer's int8 GEMMs (second forward of a 2-layer stack, eager launches)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=${1:-r02c}; pat=${2:-gemm_i8}; skip=${3:-9}; cnt=${4:-5}
LELE_B200_GRAPH=0 QS_LAYERS=2 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c $cnt \
    -f -o gpurun_out/prof_$tag python tools/quick_step.py > gpurun_out/ncu_$tag.log 2>&1
tail -5 gpurun_out/ncu_$tag.log; ls -la gpurun_out/prof_$tag.ncu-rep
