#!/bin/bash
# round 2, call 36: fast backward kernels (attention dst / src passes in the chunk layout, vectorised GELU, chunked LayerNorm backward), no fp32
# copies in LinearFn: backward suite, training-step breakdown, head-to-head training step
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_graphconv_fused.py -x -q > gpurun_out/r2/c36_tests_backward.log 2>&1
tail -4 gpurun_out/r2/c36_tests_backward.log
timeout 600 python profiles/train_breakdown.py > gpurun_out/r2/c36_train_breakdown.jsonl 2> gpurun_out/r2/c36_train_breakdown.err
cut -c1-3500 gpurun_out/r2/c36_train_breakdown.jsonl; tail -3 gpurun_out/r2/c36_train_breakdown.err
