#!/bin/bash
# round 2, call 42: cfg4-width processor parity against the unmodified reference on one GPU; bench line with the reference_gpu context leg
set -x
mkdir -p gpurun_out/r2
timeout 900 python profiles/parity_cfg4_processor.py > gpurun_out/r2/c42_parity_cfg4_processor.json 2> gpurun_out/r2/c42_parity_cfg4_processor.err
cat gpurun_out/r2/c42_parity_cfg4_processor.json; tail -4 gpurun_out/r2/c42_parity_cfg4_processor.err
timeout 900 python bench.py > gpurun_out/r2/c42_bench_cfg2_default.json 2> gpurun_out/r2/c42_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c42_bench_cfg2_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['reference_gpu'], d['parity'])" || tail -5 gpurun_out/r2/c42_bench_cfg2_default.err
