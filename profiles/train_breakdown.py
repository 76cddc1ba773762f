"""Where does one TRAINING step (forward + backward, bf16 autocast) of the cfg2 stack spend its time?  Per C-ABI entry point (CUDA events
around every launch, ops.start_timing) plus the total, so that what is left is PyTorch's share (dW GEMMs, index_select, cat, adds).
    python profiles/train_breakdown.py [--workload cfg2]"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from anemoi_core_b200 import ops  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402

wl = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "cfg2"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda")
gr = build_graph(w["grid"], w["mesh_level"])
model = bench.build_model(w, gr).to(dev).train()
x_grid, x_mesh = bench.make_inputs(w, gr)
grd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in gr.items()}
xg, xm = x_grid.to(dev), x_mesh.to(dev)
wgt = torch.randn(gr["n_grid"], w["out_grid"], generator=torch.Generator().manual_seed(7)).to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = model(xg, xm, grd)
    (y.float() * wgt).sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
step()
b.record()
torch.cuda.synchronize()
total = a.elapsed_time(b)
rec = ops.start_timing()
step()
torch.cuda.synchronize()
rec = ops.stop_timing()
tot, cnt = collections.Counter(), collections.Counter()
for name, e0, e1, fl, by in rec:
    tot[name] += e0.elapsed_time(e1)
    cnt[name] += 1
print(json.dumps({"workload": wl, "train_step_ms": round(total, 2), "c_abi_ms": round(sum(tot.values()), 2),
                  "by_entry_point": {k: {"ms": round(v, 2), "launches": cnt[k]} for k, v in tot.most_common()}}))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:25]
print(json.dumps({"top_cuda_kernels_ms": [(r.key[:90], round(r.device_time_total / 1e3, 2), r.count) for r in rows]}))
