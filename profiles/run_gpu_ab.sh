#!/bin/bash
# A/B of GEMM variants (built by `make BUILD=.. LIB=.. EXTRA=..`): correctness of the default, then isolated timings of each
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "linear or gemm" > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/ab_tests.log
echo "== default"; timeout 300 python profiles/bench_kernels.py gemm ${CUBLAS:-} --reps 20 2>&1 | tee gpurun_out/ab_kernels_default.jsonl | cut -c8-200
for v in $(ls anemoi_core_b200/lib/variants/ 2>/dev/null | sed 's/\.so$//'); do
  export ANEMOI_B200_LIB=$PWD/anemoi_core_b200/lib/variants/$v.so
  echo "== $v"; timeout 300 python profiles/bench_kernels.py gemm --reps 20 2>&1 | tee gpurun_out/ab_kernels_$v.jsonl | cut -c8-150
done
unset ANEMOI_B200_LIB
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bench_default.json 2> gpurun_out/ab_bench_default.err; cut -c1-200 gpurun_out/ab_bench_default.json
