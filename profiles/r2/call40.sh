#!/bin/bash
# round 2, call 40: breakdown of one cfg3 (GNN 16 x 1024) training step
set -x
mkdir -p gpurun_out/r2
timeout 900 python profiles/train_breakdown.py --workload cfg3 > gpurun_out/r2/c40_train_breakdown_cfg3.jsonl 2> gpurun_out/r2/c40_train_breakdown_cfg3.err
cut -c1-3800 gpurun_out/r2/c40_train_breakdown_cfg3.jsonl; tail -3 gpurun_out/r2/c40_train_breakdown_cfg3.err
