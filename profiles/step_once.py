"""One steady-state EAGER step of a bench workload between cudaProfilerStart / Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python profiles/step_once.py [cfg2]
(the launch list of exactly one step: warm-up, weight packing and CSR builds stay outside the profiled range) and for
    ncu --profile-from-start off --set full --import-source on -k regex:<kernel> -o prof python profiles/step_once.py [cfg2]
(the kernels as they run inside the step: real operands, real shapes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from anemoi_core_b200.layers._functional import freeze_packed_weights  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
gr = build_graph(w["grid"], w["mesh_level"])
model = bench.build_model(w, gr).cuda()
gd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in gr.items()}
xg, xm = (t.cuda() for t in bench.make_inputs(w, gr))
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        model(xg, xm, gd)
    freeze_packed_weights(model)
    model(xg, xm, gd)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    model(xg, xm, gd)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
