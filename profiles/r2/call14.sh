#!/bin/bash
# round 2, call 14: the one-kernel GraphConv (graphconv_fused.cu): parity tests, then timing against the decomposed form
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_graphconv_fused.py -x -q > gpurun_out/r2/c14_tests_gcf.log 2>&1
tail -15 gpurun_out/r2/c14_tests_gcf.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gnn" > gpurun_out/r2/c14_tests_gnn_parity.log 2>&1
tail -5 gpurun_out/r2/c14_tests_gnn_parity.log
timeout 600 python profiles/bench_kernels.py gcf --reps 10 > gpurun_out/r2/c14_kernels_gcf.jsonl 2>&1
cat gpurun_out/r2/c14_kernels_gcf.jsonl | cut -c1-400
