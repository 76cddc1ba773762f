#!/bin/bash
# round 2, call 38 (2 GPUs): the 2-GPU file after the training-path changes (fast backward kernels, split), head-to-head training step with the
# floored gradient metric
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s > gpurun_out/r2/c38_tests_multi.log 2>&1
tail -4 gpurun_out/r2/c38_tests_multi.log | cut -c1-2500
timeout 900 python profiles/bench_reference_gpu.py --workload cfg2 --steps 6 --train > gpurun_out/r2/c38_reference_gpu_train.json 2> gpurun_out/r2/c38_reference_gpu_train.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c38_reference_gpu_train.json')); print({k:v for k,v in d.items() if 'train' in k or k.startswith('ours')})"
