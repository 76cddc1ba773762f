#!/bin/bash
# round 2, call 35: breakdown of one cfg2 training step (ours)
set -x
mkdir -p gpurun_out/r2
timeout 600 python profiles/train_breakdown.py > gpurun_out/r2/c35_train_breakdown.jsonl 2> gpurun_out/r2/c35_train_breakdown.err
cat gpurun_out/r2/c35_train_breakdown.jsonl | cut -c1-4000; tail -5 gpurun_out/r2/c35_train_breakdown.err
