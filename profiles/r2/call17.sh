#!/bin/bash
# round 2, call 17: fused GraphConv v4 (bias in the accumulator init): full single-GPU test suite, timing with the decomposed form beside it
set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c17_tests_gpu_all.log 2>&1
tail -5 gpurun_out/r2/c17_tests_gpu_all.log
timeout 600 python profiles/bench_kernels.py gcf --reps 10 > gpurun_out/r2/c17_kernels_gcf.jsonl 2>&1
cat gpurun_out/r2/c17_kernels_gcf.jsonl | cut -c1-330
