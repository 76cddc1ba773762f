#!/bin/bash
# round 2, call 51: cost of the gather-add epilogue of GraphConv's first edge GEMM at cfg3 shapes
set -x
mkdir -p gpurun_out/r2
timeout 300 python profiles/bench_kernels.py gather --reps 15 > gpurun_out/r2/c51_gather_epilogue.jsonl 2>&1
cat gpurun_out/r2/c51_gather_epilogue.jsonl
