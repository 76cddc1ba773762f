#!/bin/bash
# round 2, call 55 (2 GPUs): HEAD multi-GPU sanity: the 2-GPU test file and the N = 2 bench line
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c55_tests_multi.log 2>&1
tail -3 gpurun_out/r2/c55_tests_multi.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2/c55_bench_cfg2_2gpu.json 2> gpurun_out/r2/c55_bench_cfg2_2gpu.err
grep '^{' gpurun_out/r2/c55_bench_cfg2_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['e2e']['value'], d.get('parity'))" || tail -5 gpurun_out/r2/c55_bench_cfg2_2gpu.err
