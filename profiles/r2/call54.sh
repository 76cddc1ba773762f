#!/bin/bash
# round 2, call 54: HEAD after the gather-table change: cfg2 default bench line (GT path must be unaffected), cfg3 line, smoke
set -x
mkdir -p gpurun_out/r2
timeout 900 python bench.py > gpurun_out/r2/c54_bench_cfg2_default.json 2> gpurun_out/r2/c54_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c54_bench_cfg2_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['reference_gpu'].get('value'), d['parity'])" || tail -5 gpurun_out/r2/c54_bench_cfg2_default.err
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c54_bench_cfg3.json 2> gpurun_out/r2/c54_bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c54_bench_cfg3.json')); print('cfg3', d['value'], d['e2e']['value'], d['roofline']['frac'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
