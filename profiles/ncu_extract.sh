#!/bin/bash
# ncu_extract.sh REPORT.ncu-rep OUTPREFIX: raw-page summary CSV (selected metrics of every captured launch) + SASS summaries; used on the GPU box so
# that only the small text artefacts travel back (gpurun merges at most 64 MiB).
rep=$1; out=$2
ncu -i "$rep" --page raw --csv 2>/dev/null | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
keep=[i for i,h in enumerate(hdr) if h in ('ID','Kernel Name','gpu__time_duration.sum','sm__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or ('issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h) or ('pipe_tensor' in h and ('pct' in h or h.endswith('.sum')))]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep if i < len(r)])
" > "${out}_raw_summary.csv"
