#!/bin/bash
# round 2, call 12: A/B of the warp-per-node attention kernel's ring depth / occupancy / batch (variant libraries, same call)
set -x
mkdir -p gpurun_out/r2
echo "default (8 slots, 3 CTAs, batch 3)" > gpurun_out/r2/c12_ab_attention.txt
timeout 200 python profiles/bench_kernels.py attn --reps 30 >> gpurun_out/r2/c12_ab_attention.txt 2>&1
for v in s6b4 s5b4 s10b2 s8b3p2 s6b4p2 s7b3p6; do
  echo "variant $v" >> gpurun_out/r2/c12_ab_attention.txt
  ANEMOI_B200_LIB=anemoi_core_b200/lib/variants/attn_$v.so timeout 200 python profiles/bench_kernels.py attn --reps 30 >> gpurun_out/r2/c12_ab_attention.txt 2>&1
done
echo "default again" >> gpurun_out/r2/c12_ab_attention.txt
timeout 200 python profiles/bench_kernels.py attn --reps 30 >> gpurun_out/r2/c12_ab_attention.txt 2>&1
grep -E "variant|default|us_median" gpurun_out/r2/c12_ab_attention.txt | cut -c1-160
