#!/bin/bash
# round-2 profile set -> gpurun_out/: ncu launch list of one forward (+ DRAM bytes), ncu full of the layer kernels, ncu launch list + full of the conv kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
tag=r02
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^(?!prep_).*" -c 720 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-exact-mode > gpurun_out/launches_$tag.log 2>&1
QS_LAYERS=2 LELE_B200_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attn_tc_kernel|gemm_i8_tc_kernel|ln_quant_cluster|quantize_rows_reg|fbank_lfr|fsmn_vt" -c 16 -f -o gpurun_out/${tag}_layer \
    python tools/quick_step.py > gpurun_out/${tag}_layer.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_tc_pixel_kernel" -s 26 -c 6 -f -o gpurun_out/${tag}_conv \
    python tools/conv_microbench.py > gpurun_out/${tag}_conv.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$tag.csv; tail -2 gpurun_out/${tag}_layer.log gpurun_out/${tag}_conv.log
