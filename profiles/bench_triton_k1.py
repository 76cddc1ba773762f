"""Head-to-head on the same B200: the reference's Triton GraphTransformer attention (`anemoi::graph_transformer_attention`,
triton/gt.py:81-179, 390-428 — the UNMODIFIED reference from baseline/_ref, imported with oracle/standins for the three packages
this image lacks) against `anemoi_b200_gt_attention_fwd` on the cfg2 shapes (ico-6 multi-scale mesh, 40 962 nodes, 327 600 edges,
H = 16, Ch = 32).

What is timed, per side, with CUDA events, L2 flushed before every launch, median of N:
  reference : lin_edge GEMM (cuBLAS, writes e [E, H*Ch]) + the custom op (Triton K1 + its fp32 -> q.dtype cast), and K1 alone
  ours      : gt_attention (lin_edge folded: raw attributes in, abar out; the self term added in the same kernel)
Also prints max |ours - reference| of the attention output on identical inputs (fp32 and bf16).
TEST / MEASUREMENT INFRASTRUCTURE: nothing under anemoi_core_b200/ imports this.
    python profiles/bench_triton_k1.py [--reps 20]
"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/models/src" if os.path.isdir("/root/reference/models/src") else os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "standins"))
import torch  # noqa: E402

from anemoi_core_b200 import ops  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 20
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1), round(min(ts), 1)


from anemoi.models.triton.gt import _gt_fwd  # noqa: E402
from anemoi.models.triton.gt import graph_transformer_attention  # noqa: E402
import triton.language as tl  # noqa: E402

gr = build_graph("o96", 6)
N, E, H, Ch, D = gr["n_mesh"], gr["proc_index"].shape[1], 16, 32, 11
C = H * Ch
ei = gr["proc_index"].to(dev)
csr = ops.build_csr(ei, N, N)
row, colptr = ei[0].contiguous(), csr.colptr
g = torch.Generator().manual_seed(0)
for dt in (torch.bfloat16, torch.float32):
    q, k, v = (torch.randn(N, C, generator=g).to(dt).to(dev) for _ in range(3))
    attr = gr["proc_attr"].to(dev)
    w_e = (torch.randn(C, D, generator=g) / D**0.5).to(dev)
    b_e = torch.randn(C, generator=g).to(dev) * 0.1
    lin = torch.nn.Linear(D, C).to(dev)
    with torch.no_grad():
        lin.weight.copy_(w_e)
        lin.bias.copy_(b_e)

        def ref_lin_edge():
            with torch.autocast("cuda", dtype=dt, enabled=dt == torch.bfloat16):
                return lin(attr)

        e = ref_lin_edge().to(dt)
        q3, k3, v3, e3 = q.view(N, H, Ch), k.view(N, H, Ch), v.view(N, H, Ch), e.view(E, H, Ch)
        dummy = torch.empty(0, dtype=torch.int64, device=dev)

        def ref_op():
            return graph_transformer_attention(q3, k3, v3, e3, row, colptr, dummy, dummy, dummy)[0]

        out_saved = torch.empty((N, H, Ch), device=dev, dtype=torch.float32)
        m = torch.empty((N, H), device=dev, dtype=torch.float32)

        def ref_k1():
            _gt_fwd[(N,)](q3, k3, v3, e3, m, row, colptr, out_saved, N, H, Ch, tl.float32)

        def ref_full():
            ee = ref_lin_edge().to(dt).view(E, H, Ch)
            return graph_transformer_attention(q3, k3, v3, ee, row, colptr, dummy, dummy, dummy)[0]

        ref_out = ref_op().reshape(N, C).float()
        # ours, in-kernel projection form (any dtype) and, for bf16, the folded form the processor uses
        attr_p = torch.zeros(E, 16, device=dev)
        attr_p[:, :D] = attr
        attr12 = attr_p[:, :12].contiguous()
        ours = ops.gt_attention(q, k, v, csr, H, edge_attr=attr12, w_edge=w_e.contiguous(), b_edge=b_e.contiguous())
        err = (ours.float() - ref_out).abs().max().item()
        rec = {"dtype": str(dt).split(".")[-1], "N": N, "E": E, "H": H, "Ch": Ch,
               "max_abs_err_vs_reference_triton": err, "ref_out_absmax": ref_out.abs().max().item()}
        rec["reference_triton_k1_us"] = timeit(ref_k1)
        rec["reference_custom_op_us(k1+cast)"] = timeit(ref_op)
        rec["reference_lin_edge_plus_op_us"] = timeit(ref_full)
        rec["ours_w_edge_form_us"] = timeit(lambda: ops.gt_attention(q, k, v, csr, H, edge_attr=attr12, w_edge=w_e, b_edge=b_e))
        if dt == torch.bfloat16:
            dp = 12
            # folded operands: qw = W_e,h^T q_h per head (what the q GEMM emits as extra columns)
            qw = torch.einsum("nhc,hca->nha", q.view(N, H, Ch).float(), w_e.view(H, Ch, D)).to(dt)
            qwp = torch.zeros(N, H, dp, dtype=dt, device=dev)
            qwp[:, :, :D] = qw
            qwp = qwp.view(N, H * dp)
            out = torch.empty(N, C + H * dp, dtype=dt, device=dev)
            x_r = torch.zeros(N, C, dtype=dt, device=dev)

            def ours_folded():
                ops.gt_attention(q, k, v, csr, H, edge_attr=attr_p, b_edge=b_e, qw=qwp, abar=out[:, C:], dp=dp, add=x_r, out=out[:, :C])

            ours_folded()
            abar = out[:, C:].float().view(N, H, dp)[:, :, :D]
            full = out[:, :C].float() + torch.einsum("nha,hca->nhc", abar, w_e.view(H, Ch, D)).reshape(N, C)
            rec["folded_max_abs_err_vs_reference_triton"] = (full - ref_out).abs().max().item()
            rec["ours_folded_us"] = timeit(ours_folded)
        print(json.dumps(rec), flush=True)
