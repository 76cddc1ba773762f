#!/bin/bash
# round 2, call 48: backward suite after the GELU-forward change; ncu --set full of the backward kernels, one launch each
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q > gpurun_out/r2/c48_tests_backward.log 2>&1; tail -2 gpurun_out/r2/c48_tests_backward.log
for k in bwd_dst_fast bwd_src_fast layer_norm_bwd_fast gelu_vec; do
  timeout 600 ncu --set full --clock-control none -k regex:$k -s 2 -c 2 -f -o /tmp/$k python profiles/train_breakdown.py > gpurun_out/r2/c48_ncu_$k.log 2>&1
  bash profiles/ncu_extract.sh /tmp/$k.ncu-rep gpurun_out/r2/c48_ncu_$k
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/r2/c48_ncu_*_raw_summary.csv")):
    rows = list(csv.reader(open(f)))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print(d.get("Kernel Name", "")[:60], "us", d.get("gpu__time_duration.sum"), "dram MB r/w", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"),
              "dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "issue%", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "lts hit", d.get("lts__t_sector_hit_rate.pct"))
PY
timeout 600 python profiles/train_breakdown.py 2>/dev/null | head -1 | cut -c1-600
