#!/bin/bash
# round 2, call 44 (2 GPUs): final verification of the round's HEAD: full -m gpu suite incl. the 2-GPU file, smoke(), N = 1 default bench, N = 2 bench
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2/c44_tests_gpu_all.log 2>&1
tail -4 gpurun_out/r2/c44_tests_gpu_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c44_smoke.log 2>&1; tail -2 gpurun_out/r2/c44_smoke.log
timeout 900 python bench.py > gpurun_out/r2/c44_bench_cfg2_default.json 2> gpurun_out/r2/c44_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c44_bench_cfg2_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['reference_gpu'].get('value'), d['parity'], d['clocks'])" || tail -5 gpurun_out/r2/c44_bench_cfg2_default.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2/c44_bench_cfg2_2gpu.json 2> gpurun_out/r2/c44_bench_cfg2_2gpu.err
grep '^{' gpurun_out/r2/c44_bench_cfg2_2gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['e2e']['value'], d.get('parity'))" || tail -5 gpurun_out/r2/c44_bench_cfg2_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2/c44_bench_reference_2gpu.json 2> gpurun_out/r2/c44_bench_reference_2gpu.err
cut -c1-300 gpurun_out/r2/c44_bench_reference_2gpu.json
