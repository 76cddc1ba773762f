#!/bin/bash
# round 2, call 20 (2 GPUs): PDL on the peer-exchange kernels too: 2-GPU tests, A/B of the 2-GPU step
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c20_tests_multi.log 2>&1
tail -3 gpurun_out/r2/c20_tests_multi.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c20_tests_backward.log 2>&1
tail -3 gpurun_out/r2/c20_tests_backward.log
for v in pdl1 pdl0 pdl1b pdl0b; do
  case $v in pdl0*) export ANEMOI_B200_PDL=0;; *) export ANEMOI_B200_PDL=1;; esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2/c20_bench2_$v.json 2> gpurun_out/r2/c20_bench2_$v.err
  grep '^{' gpurun_out/r2/c20_bench2_$v.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['e2e']['value'], d.get('parity'))" || tail -5 gpurun_out/r2/c20_bench2_$v.err
done
