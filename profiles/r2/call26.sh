#!/bin/bash
# round 2, call 26: the timeline (call 25) shows the K = 2048 mainloop at 656 cycles per k-block instead of 512: DRAM latency x 6 stages.
# A/B: L2 prefetch of the A operand 6 / 16 k-blocks beyond the ring (compile-time variants), kernels and step
set -x
mkdir -p gpurun_out/r2
O=gpurun_out/r2/c26_ab_l2ahead.txt
for v in base l2ahead6 l2ahead16 base2; do
  case $v in base*) unset ANEMOI_B200_LIB;; *) export ANEMOI_B200_LIB=$PWD/anemoi_core_b200/lib/variants/gemm_$v.so;; esac
  echo "variant $v" >> $O
  timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
done
for v in base l2ahead6 l2ahead16 base2; do
  case $v in base*) unset ANEMOI_B200_LIB;; *) export ANEMOI_B200_LIB=$PWD/anemoi_core_b200/lib/variants/gemm_$v.so;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c26_bench_$v.json 2> gpurun_out/r2/c26_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c26_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" | tee -a $O || tail -5 gpurun_out/r2/c26_bench_$v.err
done
grep -E "variant|us_median" $O | cut -c1-180
