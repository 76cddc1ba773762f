"""cfg4-WIDTH parity against the unmodified reference, on one GPU: the processor of BASELINE.json configs[3] — icosphere level-7
multi-scale mesh (163 842 nodes, 1 310 640 edges), GraphTransformer C = 1024, 16 heads — cut to 2 layers so that the reference (which
materialises [E, C] edge tensors) fits next to ours.  The reference modules come from baseline/_ref (+ oracle/standins) and run on the GPU
with their Triton attention backend; both sides get the same parameters and inputs.  Prints rel-L2 / max differences in fp32 and under
bf16 autocast, and the time per forward.  (The full cfg4 step only exists sharded over 8 GPUs: bench.py --gpus 8 --workload cfg4 checks
sharded == single-GPU there; this is the check of the single-GPU arithmetic against the reference at that width.)
    python profiles/parity_cfg4_processor.py [--layers 2]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from anemoi_core_b200 import synthetic as S  # noqa: E402
from anemoi_core_b200.distributed.shapes import GraphShardInfo  # noqa: E402
from anemoi_core_b200.layers import GraphTransformerProcessor  # noqa: E402
from oracle import reference_step as RS  # noqa: E402

layers = int(sys.argv[sys.argv.index("--layers") + 1]) if "--layers" in sys.argv else 2
dev = torch.device("cuda")
m, proc = S.icosphere_multiscale(7)
ei_np = S._sort_by_dst(proc[0], proc[1])
ea_np = S.edge_attributes(m, m, ei_np, 8, np.random.default_rng(42))
ei, ea = torch.from_numpy(ei_np).to(dev), torch.from_numpy(ea_np).to(dev)
n, C, H = int(m.shape[0]), 1024, 16
torch.manual_seed(1234)
ours = GraphTransformerProcessor(num_layers=layers, num_channels=C, num_chunks=1, num_heads=H, mlp_hidden_ratio=4, edge_dim=ea.shape[1]).eval()
sd = {k: v.detach().clone() for k, v in ours.state_dict().items()}
ours = ours.to(dev)
_, P, GSI, _ = RS._import_reference()
ref = P.GraphTransformerProcessor(num_layers=layers, num_channels=C, num_chunks=1, num_heads=H, mlp_hidden_ratio=4.0, edge_dim=ea.shape[1],
                                  layer_kernels=None, graph_attention_backend="triton")
ref.load_state_dict(sd, strict=True)
ref = ref.to(dev).eval()
x = torch.randn(n, C, generator=torch.Generator().manual_seed(5)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(statistics.median(ts), 3)


out = {"what": f"cfg4 processor: ico-7 mesh ({n} nodes, {ei.shape[1]} edges), GraphTransformer C={C} H={H}, {layers} layers", "device": torch.cuda.get_device_name(0)}
with torch.no_grad():
    for name, ac in (("fp32", False), ("bf16", True)):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            fo = lambda: ours(x, 1, GraphShardInfo(nodes=[n]), ea, ei)  # noqa: E731
            fr = lambda: ref(x, 1, GSI(nodes=[n], edges=None), ea, ei, None)  # noqa: E731
            yo, yr = fo().float(), fr().float()
            d = yo - yr
            out[name] = {"rel_l2_ours_vs_reference": (d.norm() / yr.norm()).item(), "max_abs_over_max_ref": (d.abs().max() / yr.abs().max()).item(),
                         "ours_ms": timed(fo), "reference_triton_ms": timed(fr)}
print(json.dumps(out))
