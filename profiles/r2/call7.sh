#!/bin/bash
# round 2, call 7: backward tests (fixed scale floor), parity-bar tests, kernel tests touched by the new row-statistics format, bench cfg2
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_bars.py -q > gpurun_out/r2/c7_tests_new.log 2>&1
tail -30 gpurun_out/r2/c7_tests_new.log
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_backward.py --deselect tests/test_gpu_bars.py > gpurun_out/r2/c7_tests_all.log 2>&1
tail -8 gpurun_out/r2/c7_tests_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c7_bench_cfg2.json 2> gpurun_out/r2/c7_bench_cfg2.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c7_bench_cfg2.json')); print(d['value'], d['e2e']['value'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})"
