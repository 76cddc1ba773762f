#!/bin/bash
# round 2, call 45: tail wave of the narrow GEMMs on a forked stream: tests, same-call A/B (kernels + step)
set -x
mkdir -p gpurun_out/r2
timeout 1800 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_bars.py -x -q > gpurun_out/r2/c45_tests.log 2>&1
tail -4 gpurun_out/r2/c45_tests.log
O=gpurun_out/r2/c45_ab_tail_split.txt
for v in split1 split0 split1b; do
  case $v in split0*) export ANEMOI_B200_TAIL_SPLIT=0;; *) export ANEMOI_B200_TAIL_SPLIT=1;; esac
  echo "variant $v" >> $O
  timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
done
for v in split1 split0 split1b split0b; do
  case $v in split0*) export ANEMOI_B200_TAIL_SPLIT=0;; *) export ANEMOI_B200_TAIL_SPLIT=1;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2/c45_bench_$v.json 2> gpurun_out/r2/c45_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c45_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], d['parity'], d['launches_per_step'])" | tee -a $O || tail -5 gpurun_out/r2/c45_bench_$v.err
done
grep -E "variant|us_median" $O | cut -c1-180
