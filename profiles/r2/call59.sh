#!/bin/bash
# round 2, call 59 (1 GPU): training path - residual adds in the GEMM epilogue (ANEMOI_B200_TRAIN_FUSE_RES) and dgamma / dbeta through the
# column-sum kernel (ANEMOI_B200_LN_BWD_COL_SUM): whole -m gpu suite with both ON, training-step breakdown OFF vs ON
set -x
mkdir -p gpurun_out/r2
ANEMOI_B200_TRAIN_FUSE_RES=1 ANEMOI_B200_LN_BWD_COL_SUM=1 timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c59_tests_gpu_all_on.log 2>&1; tail -3 gpurun_out/r2/c59_tests_gpu_all_on.log | cut -c1-1500
timeout 40 python profiles/train_breakdown.py > gpurun_out/r2/c59_train_breakdown_off.jsonl 2> gpurun_out/r2/c59_train_breakdown_off.err; head -1 gpurun_out/r2/c59_train_breakdown_off.jsonl | cut -c1-900
ANEMOI_B200_TRAIN_FUSE_RES=1 ANEMOI_B200_LN_BWD_COL_SUM=1 timeout 40 python profiles/train_breakdown.py > gpurun_out/r2/c59_train_breakdown_on.jsonl 2> gpurun_out/r2/c59_train_breakdown_on.err; head -1 gpurun_out/r2/c59_train_breakdown_on.jsonl | cut -c1-900
