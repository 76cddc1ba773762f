#!/bin/bash
# round 2, call 5 (8 GPUs): cfg2 strong scaling at N=8 with the peer-memory exchange inside one CUDA graph; cfg4 (O1280 -> ico-7, GT 16x1024) at N=8
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2/c5_bench_cfg2_8gpu.json 2> gpurun_out/r2/c5_bench_cfg2_8gpu.err
tail -c 1500 gpurun_out/r2/c5_bench_cfg2_8gpu.json; tail -5 gpurun_out/r2/c5_bench_cfg2_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r2/c5_bench_cfg4_8gpu.json 2> gpurun_out/r2/c5_bench_cfg4_8gpu.err
tail -c 2500 gpurun_out/r2/c5_bench_cfg4_8gpu.json; tail -15 gpurun_out/r2/c5_bench_cfg4_8gpu.err
nvidia-smi --query-gpu=index,memory.used --format=csv | head -10
