#!/bin/bash
# round 2, call 13: A/B of the residual L2 prefetch (cp.async.bulk.prefetch.tensor of the next tile's residual rows) in the RES GEMMs, same call
set -x
mkdir -p gpurun_out/r2
O=gpurun_out/r2/c13_ab_res_l2.txt
echo "default (RES_L2_AHEAD=1)" > $O
timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
echo "variant no_res_l2" >> $O
ANEMOI_B200_LIB=anemoi_core_b200/lib/variants/gemm_no_res_l2.so timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
echo "default again" >> $O
timeout 300 python profiles/bench_kernels.py gemm --reps 30 >> $O 2>&1
for v in default no_res_l2 default2 no_res_l2_2; do
  case $v in no_res_l2*) export ANEMOI_B200_LIB=anemoi_core_b200/lib/variants/gemm_no_res_l2.so;; *) unset ANEMOI_B200_LIB;; esac
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2/c13_bench_$v.json 2> gpurun_out/r2/c13_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2/c13_bench_$v.json')); print('$v', d['value'], d['e2e']['value'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" | tee -a $O
done
grep -E "variant|default|us_median" $O | cut -c1-200
