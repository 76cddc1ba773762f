#!/bin/bash
# round 2, call 56 (8 GPUs): HEAD, cfg2 at N = 8
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/c56_bench_cfg2_8gpu.json 2> gpurun_out/r2/c56_bench_cfg2_8gpu.err
grep '^{' gpurun_out/r2/c56_bench_cfg2_8gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('8gpu', d['value'], d['e2e']['value'], d.get('parity'), d['roofline']['frac'])" || tail -5 gpurun_out/r2/c56_bench_cfg2_8gpu.err
