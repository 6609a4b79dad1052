#!/bin/bash
# GPU box: the round's measurement set -> gpurun_out/ (bench line, reference arm, per-op table, ncu launch list, ncu full of the layer kernels)
tag=${1:-r01j}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
python bench.py --ops > gpurun_out/ops_$tag.jsonl 2> gpurun_out/ops_$tag.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name "regex:^(?!prep_).*" -c 720 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
QS_LAYERS=2 LELE_B200_LANES=1 ncu --set full --clock-control none --import-source on -k "regex:attn_tc_kernel|gemm_i8_tc_kernel|ln_quant_cluster|quantize_rows_reg|fbank_lfr" -c 14 -f -o gpurun_out/${tag}_layer \
    python tools/quick_step.py > gpurun_out/${tag}_layer.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json; echo; tail -c 300 gpurun_out/bench_ref_$tag.json; echo; tail -3 gpurun_out/ops_$tag.err; wc -l gpurun_out/ops_$tag.jsonl gpurun_out/launches_$tag.csv
