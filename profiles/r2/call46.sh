#!/bin/bash
# round 2, call 46: fused GraphConv with packed fp32x2 LayerNorm arithmetic: tests + timing
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_graphconv_fused.py -x -q > gpurun_out/r2/c46_tests_gcf.log 2>&1; tail -2 gpurun_out/r2/c46_tests_gcf.log
timeout 600 python profiles/bench_kernels.py gcf nodecomp --reps 10 > gpurun_out/r2/c46_kernels_gcf.jsonl 2>&1; cut -c1-260 gpurun_out/r2/c46_kernels_gcf.jsonl
