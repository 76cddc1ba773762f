#!/bin/bash
# round 2, call 25: where does the fixed cost of a GEMM launch go?  in-kernel cycle timeline (debug build)
set -x
mkdir -p gpurun_out/r2
ANEMOI_B200_LIB=anemoi_core_b200/lib/variants/gemm_timeline.so timeout 300 python profiles/gemm_timeline.py > gpurun_out/r2/c25_gemm_timeline.jsonl 2>&1
cat gpurun_out/r2/c25_gemm_timeline.jsonl
