#!/bin/bash
# round 2, call 47: ncu --set full of the new kernels of the round's second half, for the record: attention backward (dst / src fast passes),
# LayerNorm backward, vectorised GELU, segment sum (one cfg2 / cfg3-sized training step), fused GraphConv C = 32 (final)
set -x
mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bwd_dst_fast|bwd_src_fast|layer_norm_bwd_fast|gelu_vec" -c 8 -f -o /tmp/bwd python profiles/train_breakdown.py > gpurun_out/r2/c47_ncu_bwd.log 2>&1
tail -2 gpurun_out/r2/c47_ncu_bwd.log
bash profiles/ncu_extract.sh /tmp/bwd.ncu-rep gpurun_out/r2/c47_ncu_backward_kernels
GCF_C=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:graphconv_fused -c 1 -f -o /tmp/gcf python profiles/bench_kernels.py gcf nodecomp --reps 1 > gpurun_out/r2/c47_ncu_gcf.log 2>&1
bash profiles/ncu_extract.sh /tmp/gcf.ncu-rep gpurun_out/r2/c47_ncu_graphconv_fused_c32_final
python - <<'PY'
import csv
for f in ("gpurun_out/r2/c47_ncu_backward_kernels_raw_summary.csv", "gpurun_out/r2/c47_ncu_graphconv_fused_c32_final_raw_summary.csv"):
    rows = list(csv.reader(open(f)))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print(d.get("Kernel Name", "")[:70], "us", d.get("gpu__time_duration.sum"), "dram MB r/w", d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"),
              "dram%", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "issue%", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"))
PY
