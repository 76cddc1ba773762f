#!/bin/bash
# round 2, call 16: fused GraphConv v3 (division-free gather map, lane-group segmented sum): tests, timing, ncu of the C = 32 kernel
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_graphconv_fused.py -q > gpurun_out/r2/c16_tests_gcf.log 2>&1
tail -8 gpurun_out/r2/c16_tests_gcf.log
timeout 600 python profiles/bench_kernels.py gcf nodecomp --reps 10 > gpurun_out/r2/c16_kernels_gcf.jsonl 2>&1
cat gpurun_out/r2/c16_kernels_gcf.jsonl | cut -c1-330
GCF_C=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:graphconv_fused -c 2 -f -o /tmp/gcf python profiles/bench_kernels.py gcf nodecomp --reps 1 > gpurun_out/r2/c16_ncu.log 2>&1
bash profiles/ncu_extract.sh /tmp/gcf.ncu-rep gpurun_out/r2/c16_ncu_gcf32
python profiles/sass_summary.py /tmp/gcf.ncu-rep 1 > gpurun_out/r2/c16_ncu_gcf32_sass.txt 2>&1
head -60 gpurun_out/r2/c16_ncu_gcf32_sass.txt
