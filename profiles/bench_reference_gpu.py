"""GPU head-to-head of the WHOLE step on the same B200: the UNMODIFIED reference modules (baseline/_ref, with oracle/standins for
torch_geometric / hydra / anemoi.utils) run on the GPU through their own code path — PyTorch / cuBLAS kernels and, for the
GraphTransformer, the reference's Triton attention backend (its fastest, training/docs/user-guide/performance-optimisation.rst:305-312) or
the PyG backend — against this repository's step, on the cfg2 workload (BASELINE.json configs[1]) with the same parameters and inputs.

Reports ms/step (CUDA events, L2 flushed between steps, eager for the reference, eager and CUDA-graph replay for ours), and the difference
of the two outputs in bf16 autocast and in fp32.  The reference number is CONTEXT for the bench line (bench.py's reference arm is the CPU
path the task names); nothing here is imported by the package.
    python profiles/bench_reference_gpu.py [--steps 10] [--workload cfg2|small]
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402
from oracle import reference_step as RS  # noqa: E402

steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
LEG = "--bench-leg" in sys.argv  # called by bench.py (in a subprocess, with a timeout): bf16 only, the reference's fastest backend only
wl = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "cfg2"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda")
gr = build_graph(w["grid"], w["mesh_level"])
model = bench.build_model(w, gr)
sds = bench.state_dicts(model)
x_grid, x_mesh = bench.make_inputs(w, gr)
grd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in gr.items()}
xg, xm = x_grid.to(dev), x_mesh.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
model = model.to(dev)


def timed(fn, n):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return round(statistics.median(ts), 3), round(min(ts), 3)


out = {"workload": f"{wl}: {w['desc']}", "device": torch.cuda.get_device_name(0)}
ours = {}
PRECS = (("bf16", torch.bfloat16),) if LEG else (("bf16", torch.bfloat16), ("fp32", None))
for name, dt in PRECS:
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt is not None):
        ours[name] = model(xg, xm, grd).float()
        out[f"ours_eager_{name}_ms"] = timed(lambda: model(xg, xm, grd), steps)[0]
for backend in ("triton", "pyg"):
    if w["kind"] != "graphtransformer" and backend == "triton":
        continue
    if LEG and backend == "pyg" and w["kind"] == "graphtransformer":
        continue
    try:
        ref = RS.ReferenceStep(w["kind"], state_dicts=sds, attention_backend=backend, **bench._ref_kwargs(w, gr)).to(dev)
        for name, dt in PRECS:
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt is not None):
                y = ref(xg, xm, grd).float()
                ms = timed(lambda: ref(xg, xm, grd), steps)[0]
            d = (ours[name] - y)
            out[f"reference_{backend}_{name}"] = {"ms_per_step_eager": ms, "rel_l2_ours_vs_reference": round((d.norm() / y.norm()).item(), 6),
                                                  "max_abs_over_max_ref": round((d.abs().max() / y.abs().max()).item(), 6)}  # fmt: skip
        del ref
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001 - report and go on with the other backend
        out[f"reference_{backend}"] = f"unavailable: {type(e).__name__}: {str(e)[:300]}"
if "--train" in sys.argv:
    # one TRAINING step (forward + backward of sum(output * w), bf16 autocast) of the same stack: our differentiable path (layers/_train.py:
    # the sm_100a kernels through autograd Functions) against the reference's autograd (its Triton attention forward / backward kernels)
    wgt = torch.randn(gr["n_grid"], w["out_grid"], generator=torch.Generator().manual_seed(7)).to(dev)

    def train_step(fn, params):
        for p in params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = fn()
        (y.float() * wgt).sum().backward()

    model.train()
    ps = [p for p in model.parameters()]
    out["ours_train_step_bf16_ms"] = timed(lambda: train_step(lambda: model(xg, xm, grd), ps), max(3, steps // 2))[0]
    g_ours = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
    model.eval()
    for backend in ("triton", "pyg"):
        try:
            ref = RS.ReferenceStep(w["kind"], state_dicts=sds, attention_backend=backend, **bench._ref_kwargs(w, gr)).to(dev)
            for m in (ref.encoder, ref.processor, ref.decoder):
                m.train()
            rps = [p for m in (ref.encoder, ref.processor, ref.decoder) for p in m.parameters()]
            ref.no_grad = False
            ms = timed(lambda: train_step(lambda: ref(xg, xm, grd), rps), max(3, steps // 2))[0]
            # per-parameter rel-L2 of the gradients, both sides bf16 autocast.  Some gradients are zero by construction (lin_key.bias: a constant
            # added to every key drops out of the softmax) and hold only rounding noise: a tensor's norm is floored at 1e-3 of the largest one.
            pairs = []
            for part in ("encoder", "processor", "decoder"):
                for n, p in getattr(ref, part).named_parameters():
                    go = g_ours.get(f"{part}.{n}")
                    if go is not None and p.grad is not None:
                        pairs.append((f"{part}.{n}", go, p.grad.float()))
            big = max(gr_.norm().item() for _, _, gr_ in pairs)
            errs = sorted((((go - gr_).norm() / max(gr_.norm().item(), 1e-3 * big)).item(), n) for n, go, gr_ in pairs)
            out[f"reference_{backend}_train_step_bf16"] = {"ms_per_step_eager": ms, "param_grads_compared": len(pairs),
                                                           "median_param_grad_rel_l2": round(errs[len(errs) // 2][0], 5),
                                                           "worst_param_grad_rel_l2": [round(errs[-1][0], 5), errs[-1][1]]}  # fmt: skip
            del ref
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[f"reference_{backend}_train"] = f"unavailable: {type(e).__name__}: {str(e)[:300]}"
print(json.dumps(out))
