#!/bin/bash
# round 2, call 58 (1 GPU, the rest of the round's budget): training path - bias gradients through the two-stage column-sum kernel
# (ANEMOI_B200_COL_SUM) and the two-MUFU GELU backward (ANEMOI_B200_GELU_BWD_FAST): whole -m gpu suite with both ON, training-step breakdown OFF vs ON
set -x
mkdir -p gpurun_out/r2
ANEMOI_B200_COL_SUM=1 ANEMOI_B200_GELU_BWD_FAST=1 timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c58_tests_gpu_all_on.log 2>&1; tail -3 gpurun_out/r2/c58_tests_gpu_all_on.log | cut -c1-1500
timeout 60 python profiles/train_breakdown.py > gpurun_out/r2/c58_train_breakdown_off.jsonl 2> gpurun_out/r2/c58_train_breakdown_off.err; head -1 gpurun_out/r2/c58_train_breakdown_off.jsonl | cut -c1-900
ANEMOI_B200_COL_SUM=1 ANEMOI_B200_GELU_BWD_FAST=1 timeout 60 python profiles/train_breakdown.py > gpurun_out/r2/c58_train_breakdown_on.jsonl 2> gpurun_out/r2/c58_train_breakdown_on.err; head -1 gpurun_out/r2/c58_train_breakdown_on.jsonl | cut -c1-900
timeout 45 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c58_tests_backward_off.log 2>&1; tail -2 gpurun_out/r2/c58_tests_backward_off.log | cut -c1-600
