#!/bin/bash
# round 2, call 2: destination-tile attention kernel: parity tests, isolated timing (natural / Hilbert order), cfg2 bench, full GPU suite
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "tiled or tile_plan" > gpurun_out/r2/c2_tests_tile.log 2>&1
tail -5 gpurun_out/r2/c2_tests_tile.log
timeout 300 python profiles/bench_kernels.py attn tile --reps 20 > gpurun_out/r2/c2_kernels_attn_tile_natural.jsonl 2>&1
timeout 300 python profiles/bench_kernels.py attn reorder tile --reps 20 > gpurun_out/r2/c2_kernels_attn_tile_hilbert.jsonl 2>&1
cat gpurun_out/r2/c2_kernels_attn_tile_*.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c2_bench_cfg2.json 2> gpurun_out/r2/c2_bench_cfg2.err
tail -c 1500 gpurun_out/r2/c2_bench_cfg2.json; tail -3 gpurun_out/r2/c2_bench_cfg2.err
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2/c2_tests_all.log 2>&1
tail -8 gpurun_out/r2/c2_tests_all.log
