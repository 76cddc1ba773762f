#!/bin/bash
# round 2, call 1: baseline of HEAD on this box + the A/Bs round 1 left unmeasured (locality reorder, <128,2> pair tile) + Triton K1 head-to-head
set -x
mkdir -p gpurun_out/r2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
ANEMOI_B200_REORDER=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/bench_reorder.json 2> gpurun_out/r2/bench_reorder.err
python profiles/bench_kernels.py attn gemm cublas --reps 20 > gpurun_out/r2/kernels_base.jsonl 2>&1
python profiles/bench_kernels.py attn reorder --reps 20 > gpurun_out/r2/kernels_attn_reorder.jsonl 2>&1
ANEMOI_B200_GEMM_PAIR_BN128=1 python profiles/bench_kernels.py gemm --reps 20 > gpurun_out/r2/kernels_bn128.jsonl 2>&1
python profiles/bench_triton_k1.py --reps 20 > gpurun_out/r2/triton_k1.jsonl 2> gpurun_out/r2/triton_k1.err
tail -3 gpurun_out/r2/*.jsonl
