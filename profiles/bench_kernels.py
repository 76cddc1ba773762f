"""Isolated kernel timings at cfg2 shapes (ico-6 mesh, C=512, H=16, bf16): attention (folded lin_edge) and the four
per-layer GEMMs.  CUDA events, L2 flushed before every timed launch, median of N.  Usage (on the GPU box):
    python profiles/bench_kernels.py [attn] [gemm] [gc] [cublas] [--reps 30]
Prints one JSON line per kernel with us, achieved GB/s or TFLOP/s and the fraction of the measured peaks."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from anemoi_core_b200 import ops  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402

reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 30
which = [a for a in sys.argv[1:] if not a.startswith("--") and not a.isdigit()] or ["attn", "gemm", "gc"]
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
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
    return statistics.median(ts), min(ts)


g = torch.Generator().manual_seed(0)
if "attn" in which:
    gr = build_graph("o96", 6)
    N, E, H, Ch, dp = gr["n_mesh"], gr["proc_index"].shape[1], 16, 32, 12
    C = H * Ch
    csr = ops.build_csr(gr["proc_index"].to(dev), N, N)
    buf = torch.randn(N, 4 * C + H * dp, generator=g).to(torch.bfloat16).to(dev)
    ea = torch.zeros(E, 16)
    ea[:, :11] = gr["proc_attr"]
    ea = ea.to(dev)
    b_e = torch.randn(C, generator=g).to(dev)
    out = torch.empty(N, C + H * dp, dtype=torch.bfloat16, device=dev)

    def attn():
        ops.gt_attention(buf[:, :C], buf[:, C : 2 * C], buf[:, 2 * C : 3 * C], csr, H, edge_attr=ea, b_edge=b_e, qw=buf[:, 4 * C :], abar=out[:, C:], dp=dp,
                         add=buf[:, 3 * C : 4 * C], out=out[:, :C])  # fmt: skip

    if "reorder" in which:  # same kernel on the locality-ordered (Hilbert) relabelling of the graph (layers/_reorder.py)
        from anemoi_core_b200.layers import _reorder as RO

        plan = RO.locality_plan(gr["proc_index"].to(dev), N, min_nodes=0)
        csr = ops.build_csr(plan.edge_index, N, N)
        ea = ea.index_select(0, plan.edge_perm).contiguous()
        print(json.dumps({"reorder": True, "reuse16_before": RO.source_reuse(gr["proc_index"]), "reuse16_after": RO.source_reuse(plan.edge_index)}))
    if "tile" in which:  # destination-tile tensor-core kernel on the current (natural or reordered) edge list
        tplan = ops.attention_tiles(csr)
        print(json.dumps({"tile_plan": True, "n_tiles": tplan.n_tiles, "n_slots": tplan.n_slots, "reuse": round(tplan.reuse, 3)}))

        def attn():  # noqa: F811
            ops.gt_attention(buf[:, :C], buf[:, C : 2 * C], buf[:, 2 * C : 3 * C], csr, H, edge_attr=ea, b_edge=b_e, qw=buf[:, 4 * C :], abar=out[:, C:], dp=dp,
                             add=buf[:, 3 * C : 4 * C], out=out[:, :C], tiles=tplan)  # fmt: skip

    med, mn = timeit(attn)
    # algorithmic bytes: q, k, v, self read + out written (N*C*2 each) + qw/abar + per edge: src id 4 B + 64 B attributes
    alg = 5 * N * C * 2 + 2 * N * H * dp * 2 + E * (4 + 64) + 4 * N
    print(json.dumps({"kernel": "gt_attention(folded)", "us_median": round(med, 1), "us_min": round(mn, 1), "alg_MB": round(alg / 1e6, 1),
                      "GBs": round(alg / med / 1e3, 1), "frac_hbm_measured": round(alg / med / 1e3 / pk["hbm_gbs"], 3),
                      "gathered_MB": round((tplan.n_slots if "tile" in which else E) * 2 * C * 2 / 1e6, 1), "variant": "+".join(w for w in which if w in ("reorder", "tile")) or "pipe"}))  # fmt: skip

if "gemm" in which:
    M = 40962
    for name, N_, K, gelu, res in (("qkv+self+qw", 2240, 512, False, False), ("projection", 512, 704, False, True), ("mlp1+gelu", 2048, 512, True, False),
                                   ("mlp2+res", 512, 2048, False, True)):  # fmt: skip
        a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(N_, K, generator=g) / K**0.5).to(torch.bfloat16).to(dev)
        bias = torch.randn(N_, generator=g).to(dev)
        r = torch.randn(M, N_, generator=g).to(torch.bfloat16).to(dev) if res else None
        o = torch.empty(M, N_, dtype=torch.bfloat16, device=dev)
        med, mn = timeit(lambda: ops.linear(a, w, bias, gelu=gelu, residual=r, out=o))
        fl = 2.0 * M * N_ * K
        rec = {"kernel": f"linear {name} [{M}x{K}]x[{K}x{N_}]", "us_median": round(med, 1), "us_min": round(mn, 1),
               "TFLOPs": round(fl / med / 1e6, 1), "frac_tensor_burst": round(fl / med / 1e6 / pk["bf16_tflops"], 3)}
        if "cublas" in which:  # library yardstick: the bare GEMM of the same shape (no bias / GELU / residual), cuBLAS through torch.matmul
            wt = w.t()
            cm, _ = timeit(lambda: torch.matmul(a, wt, out=o))
            rec["cublas_plain_gemm_us"] = round(cm, 1)
            rec["cublas_TFLOPs"] = round(fl / cm / 1e6, 1)
        print(json.dumps(rec))

if "gc" in which:
    gr = build_graph("o96", 6)
    N, E = gr["n_mesh"], gr["proc_index"].shape[1]
    csr = ops.build_csr(gr["proc_index"].to(dev), N, N)
    for C in (1024, 512, 256):
        h = torch.randn(E, C, generator=g).to(torch.bfloat16).to(dev)
        e = torch.randn(E, C, generator=g).to(torch.bfloat16).to(dev)
        w, b = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
        out = torch.empty(N, C, dtype=torch.bfloat16, device=dev)
        med, mn = timeit(lambda: ops.graphconv_ln_aggregate(h, w, b, e, csr, out=out))
        alg = 3 * E * C * 2 + N * C * 2 + 4 * N
        print(json.dumps({"kernel": f"graphconv_ln_aggregate C={C} (E={E})", "us_median": round(med, 1), "us_min": round(mn, 1), "alg_MB": round(alg / 1e6, 1),
                          "GBs": round(alg / med / 1e3, 1), "frac_hbm_measured": round(alg / med / 1e3 / pk["hbm_gbs"], 3)}))  # fmt: skip

if "gcf" in which:
    # The one-kernel GraphConv (csrc/graphconv_fused.cu) at the widths where the operator is HBM-bound, on a graph large enough to time:
    # 1 M nodes, 8 M dst-sorted edges with local sources (|src - dst| < 4096: the gathered rows are L2 hits, as on a mesh).  Algorithmic
    # bytes: e read + e' written (2 E C b), src / dst ids (8 E), x read once + out written (2 N C b).  The decomposed form (2 node GEMMs + 3
    # edge GEMMs + LN/aggregate tail, the path wider layers take) on the same inputs is timed beside it.
    from anemoi_core_b200.layers.conv import GraphConv
    from anemoi_core_b200.layers.utils import load_layer_kernels

    n, deg = 1 << 20, 8
    E = n * deg
    dst = torch.arange(n).repeat_interleave(deg)
    src = (dst + torch.randint(-4096, 4096, (E,), generator=g)).clamp_(0, n - 1)
    ei = torch.stack([src, dst]).to(dev)
    for C in [int(c) for c in os.environ.get("GCF_C", "16,32,64").split(",")]:
        torch.manual_seed(C)
        conv = GraphConv(C, C, layer_kernels=load_layer_kernels(None)).eval().to(dev)
        x = torch.randn(n, C, generator=g).to(torch.bfloat16).to(dev)
        e = torch.randn(E, C, generator=g).to(torch.bfloat16).to(dev)
        with torch.no_grad():
            os.environ["ANEMOI_B200_GC_FUSED"] = "1"
            med, mn = timeit(lambda: conv(x, e, ei))
            os.environ["ANEMOI_B200_GC_FUSED"] = "0"
            dmed, _ = timeit(lambda: conv(x, e, ei)) if "nodecomp" not in which else (0.0, 0.0)
            os.environ["ANEMOI_B200_GC_FUSED"] = "1"
        alg = 2 * E * C * 2 + 8 * E + 2 * n * C * 2 + 4 * n
        print(json.dumps({"kernel": f"graphconv_fused C={C} (N={n}, E={E}, 3 layers, bf16)", "us_median": round(med, 1), "us_min": round(mn, 1),
                          "alg_MB": round(alg / 1e6, 1), "GBs": round(alg / med / 1e3, 1), "frac_hbm_measured": round(alg / med / 1e3 / pk["hbm_gbs"], 3),
                          "frac_hbm_8TBs": round(alg / med / 1e3 / 8000.0, 3), "GFLOP": round(10.0 * C * C * E / 1e9, 1),
                          "TFLOPs": round(10.0 * C * C * E / med / 1e6, 1), "decomposed_us": round(dmed, 1)}))  # fmt: skip
        del x, e, conv

if "l2mlp" in which:
    # Is MLP-2 (K = 2048, N = 512, residual) bound by the HBM read of its A operand?  Same GEMM with the hidden tensor cold (L2 flushed
    # before the launch) and warm (MLP-1 has just written it, nothing flushed in between), for row counts whose hidden tensor fits L2.
    for M in (9472, 18944, 40962):
        x = torch.randn(M, 512, generator=g).to(torch.bfloat16).to(dev)
        w1 = (torch.randn(2048, 512, generator=g) / 512**0.5).to(torch.bfloat16).to(dev)
        w2 = (torch.randn(512, 2048, generator=g) / 2048**0.5).to(torch.bfloat16).to(dev)
        b1, b2 = torch.randn(2048, generator=g).to(dev), torch.randn(512, generator=g).to(dev)
        h = torch.empty(M, 2048, dtype=torch.bfloat16, device=dev)
        o = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)
        cold, _ = timeit(lambda: ops.linear(h, w2, b2, residual=x, out=o))
        ts = []
        for _ in range(reps):
            flush.zero_()
            ops.linear(x, w1, b1, gelu=True, out=h)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.linear(h, w2, b2, residual=x, out=o)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(json.dumps({"kernel": f"mlp2+res M={M} (hidden {M * 4096 / 1e6:.0f} MB)", "cold_us": round(cold, 1), "warm_after_mlp1_us": round(statistics.median(ts), 1),
                          "TFLOPs_cold": round(2.0 * M * 512 * 2048 / cold / 1e6, 1), "TFLOPs_warm": round(2.0 * M * 512 * 2048 / statistics.median(ts) / 1e6, 1)}))  # fmt: skip

if "gather" in which:
    # What does the gather-add epilogue (GraphConv's first edge layer: + p_i[dst] + p_j[src], fp32 tables) cost at cfg3 shapes?
    gr = build_graph("o96", 6)
    N_, E_ = gr["n_mesh"], gr["proc_index"].shape[1]
    csr = ops.build_csr(gr["proc_index"].to(dev), N_, N_)
    for C in (1024, 512):
        e = torch.randn(E_, C, generator=g).to(torch.bfloat16).to(dev)
        w = (torch.randn(C, C, generator=g) / C**0.5).to(torch.bfloat16).to(dev)
        b = torch.randn(C, generator=g).to(dev)
        p_i, p_j = torch.randn(N_, C, generator=g).to(dev), torch.randn(N_, C, generator=g).to(dev)
        o = torch.empty(E_, C, dtype=torch.bfloat16, device=dev)
        plain, _ = timeit(lambda: ops.linear(e, w, b, gelu=True, out=o))
        gath, _ = timeit(lambda: ops.linear(e, w, b, gelu=True, gather1=(p_i, csr.dst32), gather2=(p_j, csr.src32), out=o))
        p_jh = p_j.to(torch.bfloat16)
        gath_h, _ = timeit(lambda: ops.linear(e, w, b, gelu=True, gather1=(p_i, csr.dst32), gather2=(p_jh, csr.src32), out=o))
        gath_hh, _ = timeit(lambda: ops.linear(e, w, b, gelu=True, gather1=(p_i.to(torch.bfloat16), csr.dst32), gather2=(p_jh, csr.src32), out=o))
        fl = 2.0 * E_ * C * C
        print(json.dumps({"kernel": f"edge GEMM-1 [{E_}x{C}]x[{C}x{C}] + GELU", "plain_us": round(plain, 1), "with_gather_add_us": round(gath, 1), "src_table_bf16_us": round(gath_h, 1), "both_tables_bf16_us": round(gath_hh, 1),
                          "TFLOPs_plain": round(fl / plain / 1e6, 1), "TFLOPs_gather": round(fl / gath / 1e6, 1)}))
