#!/bin/bash
# round 2, call 41: GraphConv training with the gathers fused into the first edge GEMM and their gradient as deterministic segment sums:
# backward suite, cfg3 training breakdown
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c41_tests_backward.log 2>&1
tail -4 gpurun_out/r2/c41_tests_backward.log
timeout 900 python profiles/train_breakdown.py --workload cfg3 > gpurun_out/r2/c41_train_breakdown_cfg3.jsonl 2> gpurun_out/r2/c41_train_breakdown_cfg3.err
cut -c1-2200 gpurun_out/r2/c41_train_breakdown_cfg3.jsonl; tail -3 gpurun_out/r2/c41_train_breakdown_cfg3.err
