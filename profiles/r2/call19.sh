#!/bin/bash
# round 2, call 19 (2 GPUs): training of gated MLPs / ConditionalLayerNorm (1 GPU tests), heads-strategy training + everything else in the 2-GPU
# file, PDL A/B at 2 GPUs
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c19_tests_backward.log 2>&1
tail -5 gpurun_out/r2/c19_tests_backward.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c19_tests_multi.log 2>&1
tail -8 gpurun_out/r2/c19_tests_multi.log | cut -c1-1500
for v in pdl1 pdl0 pdl1b pdl0b; do
  case $v in pdl0*) export ANEMOI_B200_PDL=0;; *) export ANEMOI_B200_PDL=1;; esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2/c19_bench2_$v.json 2> gpurun_out/r2/c19_bench2_$v.err
  grep '^{' gpurun_out/r2/c19_bench2_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['e2e']['value'], d.get('parity'))" || tail -5 gpurun_out/r2/c19_bench2_$v.err
done
