#!/bin/bash
# round 2, call 24: is MLP-2 bound by the HBM read of the hidden tensor?  cold vs L2-warm A operand
set -x
mkdir -p gpurun_out/r2
timeout 300 python profiles/bench_kernels.py l2mlp --reps 20 > gpurun_out/r2/c24_l2mlp.jsonl 2>&1
cat gpurun_out/r2/c24_l2mlp.jsonl
