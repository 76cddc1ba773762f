#!/bin/bash
# round 2, call 29: the whole step against the UNMODIFIED reference run on the same GPU (its Triton and PyG attention backends)
set -x
mkdir -p gpurun_out/r2
timeout 600 python profiles/bench_reference_gpu.py --workload small --steps 5 > gpurun_out/r2/c29_reference_gpu_small.json 2> gpurun_out/r2/c29_reference_gpu_small.err
cut -c1-1500 gpurun_out/r2/c29_reference_gpu_small.json; tail -5 gpurun_out/r2/c29_reference_gpu_small.err
timeout 900 python profiles/bench_reference_gpu.py --workload cfg2 --steps 10 > gpurun_out/r2/c29_reference_gpu_cfg2.json 2> gpurun_out/r2/c29_reference_gpu_cfg2.err
cut -c1-1500 gpurun_out/r2/c29_reference_gpu_cfg2.json; tail -5 gpurun_out/r2/c29_reference_gpu_cfg2.err
