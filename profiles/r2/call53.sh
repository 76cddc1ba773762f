#!/bin/bash
# round 2, call 53: both gather tables of GraphConv in bf16: parity bars, cfg3 A/B (both bf16 / src only / both fp32)
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests/test_gpu_bars.py tests/test_gpu_parity.py tests/test_gated_mlp.py tests/test_model_glue.py -m gpu -q > gpurun_out/r2/c53_tests.log 2>&1
tail -4 gpurun_out/r2/c53_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in both srconly none; do
  case $v in both) export ANEMOI_B200_GC_PI_BF16=1 ANEMOI_B200_GC_PJ_BF16=1;; srconly) export ANEMOI_B200_GC_PI_BF16=0 ANEMOI_B200_GC_PJ_BF16=1;; *) export ANEMOI_B200_GC_PI_BF16=0 ANEMOI_B200_GC_PJ_BF16=0;; esac
  timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/c53_bench_cfg3_$v.json 2> gpurun_out/r2/c53_bench_cfg3_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c53_bench_cfg3_$v.json')); print('$v', d['value'], d['e2e']['value'], d['parity'])" || tail -5 gpurun_out/r2/c53_bench_cfg3_$v.err
done
