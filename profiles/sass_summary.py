"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp-instructions by opcode, stall samples by
reason, and the hottest instructions.  Usage: python profiles/sass_summary.py report.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
blk = blocks[which]
print(blk[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
by_op = collections.Counter()
samples_by_op = collections.Counter()
stalls = collections.Counter()
tot_inst = tot_samp = 0
hot = []
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    by_op[op] += n
    samples_by_op[op] += s
    tot_inst += n
    tot_samp += s
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stalls[h] += int(r[ix[h]] or 0)
    hot.append((s, n, src[:90]))
print(f"total warp-instructions {tot_inst:,}  samples {tot_samp:,}  static SASS lines {len(rows)-1}")
print("-- executed by opcode (top 25)")
for op, n in by_op.most_common(25):
    print(f"  {op:12s} {n:14,d} {100*n/tot_inst:5.1f}%   samples {100*samples_by_op[op]/max(tot_samp,1):5.1f}%")
print("-- stall reasons (all samples)")
for k, v in stalls.most_common(10):
    print(f"  {k:28s} {100*v/max(tot_samp,1):5.1f}%")
print("-- hottest instructions by samples")
for s, n, src in sorted(hot, reverse=True)[:25]:
    print(f"  {100*s/max(tot_samp,1):5.2f}%  exec {n:10,d}  {src}")
