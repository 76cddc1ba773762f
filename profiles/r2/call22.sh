#!/bin/bash
# round 2, call 22: wave-aligned row blocks in MLP.run (hidden tensor stays in L2): tests, same-call A/B of the cfg2 step
set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c22_tests_gpu_all.log 2>&1
tail -4 gpurun_out/r2/c22_tests_gpu_all.log
for v in chunk48 chunk0 chunk96 chunk48b chunk0b; do
  case $v in chunk0*) export ANEMOI_B200_MLP_CHUNK_MB=0;; chunk96*) export ANEMOI_B200_MLP_CHUNK_MB=96;; *) export ANEMOI_B200_MLP_CHUNK_MB=48;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c22_bench_$v.json 2> gpurun_out/r2/c22_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c22_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d.get('parity'), {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/r2/c22_bench_$v.err
done
