#!/bin/bash
# round 2, call 37: backward suite + 2-GPU-free checks after the split change; breakdown; head-to-head training step vs the reference
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c37_tests_backward.log 2>&1
tail -3 gpurun_out/r2/c37_tests_backward.log
timeout 600 python profiles/train_breakdown.py > gpurun_out/r2/c37_train_breakdown.jsonl 2> gpurun_out/r2/c37_train_breakdown.err
cut -c1-1800 gpurun_out/r2/c37_train_breakdown.jsonl
timeout 900 python profiles/bench_reference_gpu.py --workload cfg2 --steps 6 --train > gpurun_out/r2/c37_reference_gpu_train.json 2> gpurun_out/r2/c37_reference_gpu_train.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c37_reference_gpu_train.json')); print({k:v for k,v in d.items() if 'train' in k or k.startswith('ours')})"
