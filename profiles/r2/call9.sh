#!/bin/bash
# round 2, call 9: streaming row_stats kernel (tests + step), defaults after the A/B (separate statistics pass), cfg3 / cfg5 lines for the record
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "row_stats or folded_layer_norm" > gpurun_out/r2/c9_tests_rowstats.log 2>&1
tail -5 gpurun_out/r2/c9_tests_rowstats.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c9_bench_cfg2.json 2> gpurun_out/r2/c9_bench_cfg2.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c9_bench_cfg2.json')); print(d['value'], d['e2e']['value'], d['roofline'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})"
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c9_bench_cfg3.json 2> gpurun_out/r2/c9_bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c9_bench_cfg3.json')); print(d['value'], d['e2e']['value'], {k:(v['us_per_launch'],v['launches_per_step'],v['tflops'],v['gbs']) for k,v in d['kernels'].items()})"
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c9_bench_cfg5.json 2> gpurun_out/r2/c9_bench_cfg5.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c9_bench_cfg5.json')); print(d['value'], d['rollout'])"
