#!/bin/bash
# round 2, call 57 (1 GPU, the last of the round's budget): HEAD as the driver will run it - the whole -m gpu suite, smoke(), the default bench line
set -x
mkdir -p gpurun_out/r2
timeout 170 python -m pytest tests -m gpu -x -q > gpurun_out/r2/c57_tests_gpu_all.log 2>&1; tail -2 gpurun_out/r2/c57_tests_gpu_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2/c57_smoke.log 2>&1; tail -1 gpurun_out/r2/c57_smoke.log
timeout 140 python bench.py > gpurun_out/r2/c57_bench_cfg2.json 2> gpurun_out/r2/c57_bench_cfg2.err
grep '^{' gpurun_out/r2/c57_bench_cfg2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1gpu', d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['frac'], d.get('reference_gpu'))" || tail -5 gpurun_out/r2/c57_bench_cfg2.err
