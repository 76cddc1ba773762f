#!/bin/bash
# Final round-1 evidence for HEAD (v10): bench lines (cfg2 with CPU baseline, cfg3, cfg5), isolated kernels, ncu launch list, ncu --set full of the
# per-layer GEMMs / attention / graphconv tail.  The .ncu-rep files are summarised ON the box (profiles/ncu_extract.sh, sass_summary.py) and
# deleted: gpurun merges at most 64 MiB back.
set -x
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_cfg2_v10.json 2> gpurun_out/bench_cfg2_v10.err
python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3_v10.json 2> gpurun_out/bench_cfg3_v10.err
python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_v10.json 2> gpurun_out/bench_cfg5_v10.err
python profiles/bench_kernels.py attn gemm gc cublas --reps 20 > gpurun_out/kernels_v10.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_cfg2_v10.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_v10.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -c 16 -f -o /tmp/gemm_v10 python profiles/bench_kernels.py gemm --reps 1 > gpurun_out/ncu_gemm_v10.log 2>&1
bash profiles/ncu_extract.sh /tmp/gemm_v10.ncu-rep gpurun_out/ncu_gemm_v10
for k in 6 14 22 30; do python profiles/sass_summary.py /tmp/gemm_v10.ncu-rep $k > gpurun_out/ncu_gemm_v10_sass_$k.txt 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gt_attention|graphconv_ln" -c 16 -f -o /tmp/attn_gc_v10 python profiles/bench_kernels.py attn gc --reps 1 > gpurun_out/ncu_attn_v10.log 2>&1
bash profiles/ncu_extract.sh /tmp/attn_gc_v10.ncu-rep gpurun_out/ncu_attn_gc_v10
python profiles/sass_summary.py /tmp/attn_gc_v10.ncu-rep 6 > gpurun_out/ncu_attn_v10_sass.txt 2>&1
du -sh gpurun_out
