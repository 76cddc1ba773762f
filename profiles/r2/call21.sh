#!/bin/bash
# round 2, call 21: serpentine (L2-aware) traversal order of the GraphTransformer block's kernels: tests, same-call A/B of the cfg2 step
set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c21_tests_gpu_all.log 2>&1
tail -4 gpurun_out/r2/c21_tests_gpu_all.log
for v in serp1 serp0 serp1b serp0b; do
  case $v in serp0*) export ANEMOI_B200_SERPENTINE=0;; *) export ANEMOI_B200_SERPENTINE=1;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c21_bench_$v.json 2> gpurun_out/r2/c21_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c21_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d.get('parity'), {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/r2/c21_bench_$v.err
done
