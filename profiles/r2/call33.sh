#!/bin/bash
# round 2, call 33: full GPU suite with the fp32 tensor-core path
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2/c33_tests_gpu_all.log 2>&1
tail -6 gpurun_out/r2/c33_tests_gpu_all.log
