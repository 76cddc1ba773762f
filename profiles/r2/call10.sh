#!/bin/bash
# round 2, call 10 (4 GPUs): smoke(), the reference arm on the box (unmodified reference modules from baseline/_ref), default bench with CPU baseline, N=4 line
set -x
mkdir -p gpurun_out/r2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c10_smoke.log 2>&1; tail -3 gpurun_out/r2/c10_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2/c10_bench_reference.json 2> gpurun_out/r2/c10_bench_reference.err; cut -c1-900 gpurun_out/r2/c10_bench_reference.json; tail -2 gpurun_out/r2/c10_bench_reference.err
timeout 900 python bench.py > gpurun_out/r2/c10_bench_cfg2_default.json 2> gpurun_out/r2/c10_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c10_bench_cfg2_default.json')); print(d['value'], d['e2e'], d['roofline'], d['cpu_baseline'], d['parity'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2/c10_bench_cfg2_4gpu.json 2> gpurun_out/r2/c10_bench_cfg2_4gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2/c10_bench_cfg2_4gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['parity'], d['config']['launch'])"
