#!/bin/bash
# GPU call: bench line + isolated kernel timings + ncu launch list + ncu full capture of GEMM / attention kernels (HEAD = v8)
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2_v8.json 2> gpurun_out/bench_cfg2_v8.err
python profiles/bench_kernels.py attn gemm --reps 20 > gpurun_out/kernels_v8.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_cfg2_v8.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/launches_v8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -c 16 -f -o gpurun_out/gemm_v8 python profiles/bench_kernels.py gemm --reps 1 > gpurun_out/ncu_gemm_v8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gt_attention -c 4 -f -o gpurun_out/attn_v8 python profiles/bench_kernels.py attn --reps 1 > gpurun_out/ncu_attn_v8.log 2>&1
cat gpurun_out/bench_cfg2_v8.json; cat gpurun_out/kernels_v8.jsonl
