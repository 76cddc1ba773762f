#!/bin/bash
# round 2, call 50: bench line with the reference_gpu leg in a subprocess
set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py > gpurun_out/r2/c50_bench_cfg2_default.json 2> gpurun_out/r2/c50_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c50_bench_cfg2_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['reference_gpu'], d['parity'])" || tail -5 gpurun_out/r2/c50_bench_cfg2_default.err
timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c50_bench_cfg5.json 2> gpurun_out/r2/c50_bench_cfg5.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c50_bench_cfg5.json')); print('cfg5', d['value'], d['rollout'], d.get('reference_gpu'))" || tail -5 gpurun_out/r2/c50_bench_cfg5.err
