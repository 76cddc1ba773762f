#!/bin/bash
# round 2, call 52: bf16 table for the src-indexed gather term of GraphConv (gather loads hoisted in the epilogue): tests, microbench, cfg3 A/B
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c52_tests_gpu_all.log 2>&1
tail -4 gpurun_out/r2/c52_tests_gpu_all.log
timeout 300 python profiles/bench_kernels.py gather --reps 15 > gpurun_out/r2/c52_gather_epilogue.jsonl 2>&1
cat gpurun_out/r2/c52_gather_epilogue.jsonl
for v in pj1 pj0; do
  case $v in pj0*) export ANEMOI_B200_GC_PJ_BF16=0;; *) export ANEMOI_B200_GC_PJ_BF16=1;; esac
  timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/c52_bench_cfg3_$v.json 2> gpurun_out/r2/c52_bench_cfg3_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c52_bench_cfg3_$v.json')); print('$v', d['value'], d['e2e']['value'], d['parity'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" || tail -5 gpurun_out/r2/c52_bench_cfg3_$v.err
done
