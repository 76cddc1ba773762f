#!/bin/bash
# round 2, call 31 (smallest-terms-first order): fp32 GEMMs on the tensor cores (exact bf16 x 3 split): full GPU suite (every fp32 parity test now runs through it), fp32 step time
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c31_tests_gpu_all.log 2>&1
tail -6 gpurun_out/r2/c31_tests_gpu_all.log
timeout 900 python profiles/bench_reference_gpu.py --workload cfg2 --steps 10 > gpurun_out/r2/c31_reference_gpu_cfg2.json 2> gpurun_out/r2/c31_reference_gpu_cfg2.err
cut -c1-1500 gpurun_out/r2/c31_reference_gpu_cfg2.json; tail -5 gpurun_out/r2/c31_reference_gpu_cfg2.err
