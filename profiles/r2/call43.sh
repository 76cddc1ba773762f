#!/bin/bash
# round 2, call 43: epilogue with the round's shared-memory loads hoisted above its stores (the 8-column groups no longer serialise): GPU suite,
# same-call A/B against the previous library (kernels and step)
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c43_tests_gpu_all.log 2>&1
tail -4 gpurun_out/r2/c43_tests_gpu_all.log
O=gpurun_out/r2/c43_ab_epilogue_hoist.txt
for v in new before new2; do
  case $v in before*) export ANEMOI_B200_LIB=$PWD/anemoi_core_b200/lib/variants/gemm_before_hoist.so;; *) unset ANEMOI_B200_LIB;; esac
  echo "variant $v" >> $O
  timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
done
for v in new before new2 before2; do
  case $v in before*) export ANEMOI_B200_LIB=$PWD/anemoi_core_b200/lib/variants/gemm_before_hoist.so;; *) unset ANEMOI_B200_LIB;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/c43_bench_$v.json 2> gpurun_out/r2/c43_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c43_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d['roofline']['frac'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" | tee -a $O || tail -5 gpurun_out/r2/c43_bench_$v.err
done
grep -E "variant|us_median" $O | cut -c1-180
