#!/bin/bash
# round 2, call 8: same-box A/B of the fused row statistics (new shifted (mean, M2) partials), clean one-step ncu launch list, ncu --set full of the
# step's GEMM / attention launches (DRAM traffic, tensor-pipe utilisation) -> profiles/r2/traffic_cfg2.json
set -x
mkdir -p gpurun_out/r2
for v in 1 0 1 0; do
  ANEMOI_B200_FUSED_ROW_STATS=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FUSED_ROW_STATS=$v', round(d['value'],3), round(d['e2e']['value'],3), {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})" >> gpurun_out/r2/c8_ab_row_stats.txt
done
cat gpurun_out/r2/c8_ab_row_stats.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/c8_launches_cfg2_one_step.csv python profiles/step_once.py cfg2 > gpurun_out/r2/c8_launches.log 2>&1
tail -3 gpurun_out/r2/c8_launches.log; wc -l gpurun_out/r2/c8_launches_cfg2_one_step.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|gt_attention" -c 30 -f -o /tmp/step_full python profiles/step_once.py cfg2 > gpurun_out/r2/c8_ncu_full.log 2>&1
tail -3 gpurun_out/r2/c8_ncu_full.log
bash profiles/ncu_extract.sh /tmp/step_full.ncu-rep gpurun_out/r2/c8_ncu_step
python profiles/ncu_traffic.py gpurun_out/r2/c8_ncu_step_raw_summary.csv gemm_bf16_tcgen05 linear_tcgen05 cfg2 gpurun_out/r2/traffic_cfg2.json
python profiles/ncu_traffic.py gpurun_out/r2/c8_ncu_step_raw_summary.csv gt_attention gt_attention cfg2 gpurun_out/r2/traffic_cfg2_attention.json
