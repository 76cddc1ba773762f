#!/bin/bash
# round 2, call 28: final state of the round on one GPU: smoke(), default bench (with the CPU baseline), reference arm, cfg3 / cfg5 lines,
# one-step ncu launch list, in-step ncu --set full of the GEMM / attention launches -> profiles/r2/traffic_cfg2*.json
set -x
mkdir -p gpurun_out/r2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/c28_smoke.log 2>&1; tail -3 gpurun_out/r2/c28_smoke.log
timeout 900 python bench.py > gpurun_out/r2/c28_bench_cfg2_default.json 2> gpurun_out/r2/c28_bench_cfg2_default.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c28_bench_cfg2_default.json')); print(d['value'], d['e2e'], d['roofline'], d['cpu_baseline'], d['parity'], d.get('gpu_launches'), d.get('clocks'))" || tail -5 gpurun_out/r2/c28_bench_cfg2_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/c28_bench_reference.json 2> gpurun_out/r2/c28_bench_reference.err; cut -c1-700 gpurun_out/r2/c28_bench_reference.json; tail -2 gpurun_out/r2/c28_bench_reference.err
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c28_bench_cfg3.json 2> gpurun_out/r2/c28_bench_cfg3.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c28_bench_cfg3.json')); print('cfg3', d['value'], d['e2e']['value'], {k:(v['us_per_launch'],v['launches_per_step']) for k,v in d['kernels'].items()})"
timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/c28_bench_cfg5.json 2> gpurun_out/r2/c28_bench_cfg5.err
python -c "
import json; d=json.load(open('gpurun_out/r2/c28_bench_cfg5.json')); print('cfg5', d['value'], d['rollout'])"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/c28_launches_cfg2_one_step.csv python profiles/step_once.py cfg2 > gpurun_out/r2/c28_launches.log 2>&1
tail -3 gpurun_out/r2/c28_launches.log; wc -l gpurun_out/r2/c28_launches_cfg2_one_step.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_bf16_tcgen05|gt_attention" -c 30 -f -o /tmp/step_full python profiles/step_once.py cfg2 > gpurun_out/r2/c28_ncu_full.log 2>&1
tail -3 gpurun_out/r2/c28_ncu_full.log
bash profiles/ncu_extract.sh /tmp/step_full.ncu-rep gpurun_out/r2/c28_ncu_step
python profiles/ncu_traffic.py gpurun_out/r2/c28_ncu_step_raw_summary.csv gemm_bf16_tcgen05 linear_tcgen05 cfg2 gpurun_out/r2/traffic_cfg2.json
python profiles/ncu_traffic.py gpurun_out/r2/c28_ncu_step_raw_summary.csv gt_attention gt_attention cfg2 gpurun_out/r2/traffic_cfg2_attention.json
cat gpurun_out/r2/traffic_cfg2.json | cut -c1-600
