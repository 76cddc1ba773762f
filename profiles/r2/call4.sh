#!/bin/bash
# round 2, call 4 (2 GPUs): peer-memory halo exchange: NCCL-parity test, whole-graph capture, bench cfg2 at N=2 (peer vs NCCL A/B)
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c4_tests_multi.log 2>&1
tail -15 gpurun_out/r2/c4_tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c4_bench_cfg2_2gpu_peer.json 2> gpurun_out/r2/c4_bench_cfg2_2gpu_peer.err
tail -c 1200 gpurun_out/r2/c4_bench_cfg2_2gpu_peer.json; tail -5 gpurun_out/r2/c4_bench_cfg2_2gpu_peer.err
ANEMOI_B200_PEER=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2/c4_bench_cfg2_2gpu_nccl.json 2> gpurun_out/r2/c4_bench_cfg2_2gpu_nccl.err
tail -c 600 gpurun_out/r2/c4_bench_cfg2_2gpu_nccl.json; tail -5 gpurun_out/r2/c4_bench_cfg2_2gpu_nccl.err
