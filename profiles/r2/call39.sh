#!/bin/bash
# round 2, call 39: cfg3 (N320 grid -> ico-6, GNN 16 x 1024) head-to-head against the unmodified reference on the same GPU: forward and one training step
set -x
mkdir -p gpurun_out/r2
timeout 1200 python profiles/bench_reference_gpu.py --workload cfg3 --steps 4 --train > gpurun_out/r2/c39_reference_gpu_cfg3.json 2> gpurun_out/r2/c39_reference_gpu_cfg3.err
cut -c1-2500 gpurun_out/r2/c39_reference_gpu_cfg3.json; tail -8 gpurun_out/r2/c39_reference_gpu_cfg3.err
nvidia-smi --query-gpu=memory.used --format=csv
