#!/bin/bash
# round 2, call 6: backward kernels + autograd wiring against the reference gradient fixtures; full GPU suite (attention kernels gained the lse output)
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py -q > gpurun_out/r2/c6_tests_backward.log 2>&1
tail -40 gpurun_out/r2/c6_tests_backward.log
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_backward.py > gpurun_out/r2/c6_tests_all.log 2>&1
tail -8 gpurun_out/r2/c6_tests_all.log
