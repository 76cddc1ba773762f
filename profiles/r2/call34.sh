#!/bin/bash
# round 2, call 34: one TRAINING step (forward + backward, bf16 autocast) of the cfg2 stack: ours vs the unmodified reference on the same GPU
set -x
mkdir -p gpurun_out/r2
timeout 900 python profiles/bench_reference_gpu.py --workload cfg2 --steps 6 --train > gpurun_out/r2/c34_reference_gpu_train.json 2> gpurun_out/r2/c34_reference_gpu_train.err
cut -c1-2500 gpurun_out/r2/c34_reference_gpu_train.json; tail -8 gpurun_out/r2/c34_reference_gpu_train.err
