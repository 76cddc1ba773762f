#!/bin/bash
# round 2, call 23: deep residual staging (4 buffers / warp, 2 residual loads in flight, 5 operand stages) for the RES GEMMs: tests, same-call A/B
set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_bars.py -x -q > gpurun_out/r2/c23_tests.log 2>&1
tail -4 gpurun_out/r2/c23_tests.log
O=gpurun_out/r2/c23_ab_resd.txt
for v in resd1 resd0 resd1b; do
  case $v in resd0*) export ANEMOI_B200_GEMM_RESD=0;; *) export ANEMOI_B200_GEMM_RESD=1;; esac
  echo "variant $v" >> $O
  timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
done
for v in resd1 resd0 resd1b resd0b; do
  case $v in resd0*) export ANEMOI_B200_GEMM_RESD=0;; *) export ANEMOI_B200_GEMM_RESD=1;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c23_bench_$v.json 2> gpurun_out/r2/c23_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c23_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d.get('parity'), {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" | tee -a $O || tail -5 gpurun_out/r2/c23_bench_$v.err
done
grep -E "variant|us_median" $O | cut -c1-200
