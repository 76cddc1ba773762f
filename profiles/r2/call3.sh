#!/bin/bash
# round 2, call 3: tile attention v2 (coalesced slot list, attributes staged by cp.async, lane-per-edge bias): parity + timing
set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "tiled or tile_plan" > gpurun_out/r2/c3_tests_tile.log 2>&1
tail -5 gpurun_out/r2/c3_tests_tile.log
timeout 300 python profiles/bench_kernels.py attn reorder tile --reps 20 > gpurun_out/r2/c3_kernels_attn_tile_hilbert.jsonl 2>&1
timeout 300 python profiles/bench_kernels.py attn tile --reps 20 > gpurun_out/r2/c3_kernels_attn_tile_natural.jsonl 2>&1
cat gpurun_out/r2/c3_kernels_attn_tile_*.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gt_attention_tile -c 2 -f -o /tmp/attn_tile python profiles/bench_kernels.py attn reorder tile --reps 1 > gpurun_out/r2/c3_ncu.log 2>&1
bash profiles/ncu_extract.sh /tmp/attn_tile.ncu-rep gpurun_out/r2/c3_ncu_attn_tile
python profiles/sass_summary.py /tmp/attn_tile.ncu-rep 1 > gpurun_out/r2/c3_ncu_attn_tile_sass.txt 2>&1
cat gpurun_out/r2/c3_ncu_attn_tile_raw_summary.csv | head -5
