#!/bin/bash
# usage: tools/ncu_times.sh <tag> [kernel regex]   (GPU box) -> gpurun_out/times_<tag>.csv : per-launch device time of the matching kernels
tag=$1; pat=${2:-gemm_i8_tc_kernel}
QS_LAYERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$pat" -c 60 --csv --log-file gpurun_out/times_$tag.csv python tools/quick_step.py > gpurun_out/times_$tag.log 2>&1
