#!/bin/bash
# round 2, call 11 (2 GPUs): sharded TRAINING step == single-GPU training step (autograd halves of the exchange), plus the rest of the 2-GPU file
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c11_tests_multi.log 2>&1
tail -12 gpurun_out/r2/c11_tests_multi.log | cut -c1-3000
