"""ncu_traffic.py RAW_SUMMARY.csv KERNEL_REGEX KERNEL_CLASS WORKLOAD OUT.json — average DRAM bytes (read + write) per launch and tensor-pipe
utilisation of the launches whose name matches, from a `profiles/ncu_extract.sh` summary of an `ncu --set full` capture.  The JSON is what
bench.py reports as roofline.traffic for that workload (profiles/r2/traffic_<workload>.json)."""
import csv
import json
import re
import sys

path, pat, klass, workload, out = sys.argv[1:6]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] == "":
        return None
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"mbyte": 1e6, "gbyte": 1e9, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0) if "byte" in u else v


sel = [r for r in rows[2:] if len(r) > col["Kernel Name"] and re.search(pat, r[col["Kernel Name"]])]
tot = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in sel]
tkey = next((h for h in hdr if "pipe_tensor" in h and "cycles_active" in h and "pct_of_peak_sustained_active" in h), None)
tens = [val(r, tkey) for r in sel] if tkey else []
tens = [t for t in tens if t is not None]
rec = {"workload": workload, "kernel_class": klass, "kernel_regex": pat, "launches": len(sel), "avg_dram_bytes_per_launch": sum(tot) / max(len(tot), 1),
       "per_launch_dram_bytes": tot, "tensor_pipe_pct_of_peak": tens, "tensor_pipe_metric": tkey, "avg_tensor_pipe_pct": (sum(tens) / len(tens) if tens else None),
       "avg_us": sum(val(r, "gpu__time_duration.sum") for r in sel) / max(len(sel), 1), "source": path}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in rec.items() if not isinstance(v, list)}))
