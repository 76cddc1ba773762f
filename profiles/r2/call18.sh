#!/bin/bash
# round 2, call 18: programmatic dependent launch (PDL) for the GEMM / attention / row-statistics kernels: full single-GPU suite, then same-call A/B
set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c18_tests_gpu_all.log 2>&1
tail -5 gpurun_out/r2/c18_tests_gpu_all.log
for v in pdl1 pdl0 pdl1b pdl0b; do
  case $v in pdl0*) export ANEMOI_B200_PDL=0;; *) export ANEMOI_B200_PDL=1;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c18_bench_$v.json 2> gpurun_out/r2/c18_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c18_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d.get('parity'), {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/r2/c18_bench_$v.err
done
