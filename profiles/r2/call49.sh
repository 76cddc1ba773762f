#!/bin/bash
# round 2, call 49: HEAD: full single-GPU -m gpu suite + smoke + default bench line
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2/c49_tests_gpu_all.log 2>&1
tail -3 gpurun_out/r2/c49_tests_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c49_smoke.log 2>&1; tail -1 gpurun_out/r2/c49_smoke.log
timeout 900 python bench.py > gpurun_out/r2/c49_bench_cfg2_default.json 2> gpurun_out/r2/c49_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c49_bench_cfg2_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['reference_gpu'].get('value'), d['parity'], d['gpu_launches'])" || tail -5 gpurun_out/r2/c49_bench_cfg2_default.err
