#!/bin/bash
# round 2, call 60 (2 GPUs, the last seconds of the round's budget): the 2-GPU NCCL test file at HEAD (training defaults: column-sum bias gradients,
# two-MUFU GELU backward, residuals in the GEMM epilogue)
set -x
mkdir -p gpurun_out/r2
timeout 50 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c60_tests_multi.log 2>&1
tail -3 gpurun_out/r2/c60_tests_multi.log | cut -c1-400
