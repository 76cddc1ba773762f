"""Backward parity on a B200 (-m gpu), SURVEY.md §8f rank 3: the drop-in modules in training mode against gradient fixtures of the
UNMODIFIED reference modules (tests/golden/grads.pt, oracle/gen_grad_golden.py: PyTorch autograd on CPU, fp32).

Bar (written here): forward and every gradient (all parameters, node inputs, edge attributes) within 1e-4 of the tensor's own scale in
fp32 — the reference's bar for its fused attention op, forward AND backward (models/tests/integration/triton/test_triton_gt.py:135-136,
179-184: atol 1e-4) — and rel-L2 <= 5e-2 under bf16 autocast.  Kernel-level checks compare the backward kernels with PyTorch autograd of
a plain statement of the same op on the same device."""
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu


def close(a, b, what, tol=1e-4, floor=1e-6):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    scale = max(b.abs().max().item(), floor)
    err = (a - b).abs().max().item() / scale
    assert err <= tol, f"{what}: max|d| / max|ref| = {err:.3e} > {tol:g}"


def check_grads(m, golden_grads, what, tol=1e-4):
    got = {n: p.grad for n, p in m.named_parameters()}
    assert set(golden_grads["params"]) <= set(got), f"{what}: missing parameters {set(golden_grads['params']) - set(got)}"
    # Some gradients are ZERO by construction (lin_key.bias: a constant added to every key shifts all scores of a destination equally and
    # drops out of the softmax) and the reference holds 1e-7 rounding noise there: a tensor's scale is floored at 1e-3 of the largest gradient.
    floor = 1e-3 * max(g.abs().max().item() for g in golden_grads["params"].values())
    for n, g in golden_grads["params"].items():
        assert got[n] is not None, f"{what}: no gradient for {n}"
        close(got[n], g, f"{what} d{n}", tol, floor)


def shard1(n=None):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    return GraphShardInfo(nodes=None if n is None else [n], edges=None)


@pytest.mark.parametrize("name", ["gt_processor", "gt_processor_qknorm", "gnn_processor"])
def test_processor_gradients_match_reference(golden, name):
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    c = golden("grads")[name]
    m = (GNNProcessor if name.startswith("gnn") else GraphTransformerProcessor)(**c["cfg"])
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    x, ea = c["x"].cuda().requires_grad_(), c["edge_attr"].cuda().requires_grad_()
    y = m(x, 1, shard1(x.shape[0]), ea, c["edge_index"].cuda())
    assert y.requires_grad and y.dtype == torch.float32
    close(y, c["y"], f"{name} forward")
    (y * c["w"].cuda()).sum().backward()
    close(x.grad, c["grads"]["x"], f"{name} dx")
    close(ea.grad, c["grads"]["edge_attr"], f"{name} dedge_attr")
    check_grads(m, c["grads"], name)


@pytest.mark.parametrize("name", ["gt_forward_mapper", "gt_backward_mapper", "gnn_forward_mapper", "gnn_backward_mapper"])
def test_mapper_gradients_match_reference(golden, name):
    import anemoi_core_b200.layers as L
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo

    c = golden("grads")[name]
    cls = {"gt_forward_mapper": L.GraphTransformerForwardMapper, "gt_backward_mapper": L.GraphTransformerBackwardMapper,
           "gnn_forward_mapper": L.GNNForwardMapper, "gnn_backward_mapper": L.GNNBackwardMapper}[name]  # fmt: skip
    m = cls(**c["cfg"])
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    xs, xd, ea = (c[k].cuda().requires_grad_() for k in ("x_src", "x_dst", "edge_attr"))
    out = m((xs, xd), 1, BipartiteGraphShardInfo(), ea, c["edge_index"].cuda())
    if name == "gnn_forward_mapper":
        ys = [out[1], out[0]]
    elif name == "gt_forward_mapper":
        ys = [out[1]]
    else:
        ys = [out]
    for y, yr in zip(ys, c["y"]):
        close(y, yr, f"{name} forward")
    sum((y * w.cuda()).sum() for y, w in zip(ys, c["w"])).backward()
    for k, t in (("x_src", xs), ("x_dst", xd), ("edge_attr", ea)):
        if c["grads"][k] is None:
            assert t.grad is None or t.grad.abs().max().item() == 0.0
        else:
            close(t.grad, c["grads"][k], f"{name} d{k}")
    check_grads(m, c["grads"], name)


def test_attention_op_forward_backward_like_test_triton_gt(golden):
    """The reference's own op-level test (test_triton_gt.py:117-184): fused attention == PyG conv, forward and backward, atol 1e-4 —
    through the module (GraphTransformerConv) and through the registered custom op with the reference's signature."""
    import anemoi_core_b200.torch_ops  # noqa: F401  (registers torch.ops.anemoi_b200.graph_transformer_attention)
    from anemoi_core_b200.layers import GraphTransformerConv

    for c in golden("grads")["gt_conv"]:
        n_dst, h, d = c["q"].shape
        n_src = c["k"].shape[0]
        for via in ("module", "custom_op"):
            q, k, v, e = (c[n].cuda().requires_grad_() for n in ("q", "k", "v", "e"))
            ei = c["edge_index"].cuda()
            if via == "module":
                out = GraphTransformerConv(out_channels=d)(q, k, v, e, ei, size=(n_src, n_dst))
            else:
                (row, colptr), _, (rowptr, edge_ids, edge_dst) = _csc_with_reverse(c["edge_index"], n_src, n_dst)
                out, out_saved, m = torch.ops.anemoi_b200.graph_transformer_attention(q, k, v, e, row.cuda(), colptr.cuda(), rowptr.cuda(), edge_ids.cuda(),
                                                                                     edge_dst.cuda())  # fmt: skip
                assert out_saved.dtype == torch.float32 and m.shape == (n_dst, h)
                # m = log-sum-exp of the scaled scores, zeros for rows without edges (gt.py:112-119, 170-178)
                assert torch.all(m[-2:] == 0)
            torch.testing.assert_close(out.cpu(), c["out"], atol=1e-4, rtol=0)
            (out * c["w"].cuda()).sum().backward()
            for t, name in ((q, "dq"), (k, "dk"), (v, "dv"), (e, "de")):
                torch.testing.assert_close(t.grad.cpu(), c[name], atol=1e-4, rtol=0)


def _csc_with_reverse(edge_index, n_src, n_dst):
    """(row, colptr), perm, (rowptr, edge_ids, edge_dst) as triton/utils.py:25-70 returns them with reverse=True."""
    row, col = edge_index[0], edge_index[1]
    colptr = R.index2ptr(col, n_dst)
    rowptr = R.index2ptr(torch.sort(row, stable=True).values, n_src)
    edge_ids = torch.argsort(row, stable=True)
    return (row, colptr), None, (rowptr, edge_ids, col)


def test_fake_impl_and_opcheck():
    import anemoi_core_b200.torch_ops  # noqa: F401

    g = torch.Generator().manual_seed(0)
    n_src, n_dst, h, d, E = 20, 30, 4, 8, 90
    ei = torch.stack([torch.randint(0, n_src, (E,), generator=g), torch.randint(0, n_dst, (E,), generator=g)])
    ei = ei[:, torch.sort(ei[1], stable=True)[1]]
    (row, colptr), _, (rowptr, edge_ids, edge_dst) = _csc_with_reverse(ei, n_src, n_dst)
    args = [torch.randn(n_dst, h, d, generator=g).cuda(), torch.randn(n_src, h, d, generator=g).cuda(), torch.randn(n_src, h, d, generator=g).cuda(),
            torch.randn(E, h, d, generator=g).cuda(), row.cuda(), colptr.cuda(), rowptr.cuda(), edge_ids.cuda(), edge_dst.cuda()]  # fmt: skip
    torch.library.opcheck(torch.ops.anemoi_b200.graph_transformer_attention.default, args, test_utils=("test_schema", "test_faketensor"))


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_attention_backward_kernel_mid_size(dt):
    """gt_attention_bwd at C = 512 / H = 16 on an ico-4 mesh against PyTorch autograd of the oracle attention on the same device."""
    from anemoi_core_b200 import autograd as AG
    from anemoi_core_b200 import ops
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph("o32", 4)
    N, H, Ch = gr["n_mesh"], 16, 32
    C = H * Ch
    ei = gr["proc_index"].cuda()
    E = ei.shape[1]
    g = torch.Generator().manual_seed(3)
    q, k, v = (torch.randn(N, C, generator=g).to(dt).cuda().requires_grad_() for _ in range(3))
    e = (0.5 * torch.randn(E, C, generator=g)).to(dt).cuda().requires_grad_()
    w = torch.randn(N, C, generator=g).cuda()
    csr = ops.build_csr(ei, N, N)
    out = AG.gt_attention(q, k, v, e, csr, H)
    (out.float() * w).sum().backward()
    got = [t.grad.float().clone() for t in (q, k, v, e)]
    q2, k2, v2, e2 = (t.detach().float().requires_grad_() for t in (q, k, v, e))
    ref = R.gt_attention(q2.view(N, H, Ch), k2.view(N, H, Ch), v2.view(N, H, Ch), e2.view(E, H, Ch), ei, N).view(N, C)
    (ref * w).sum().backward()
    tol = 2e-4 if dt == torch.float32 else 3e-2
    close(out, ref, "attention forward", tol)
    for a, b, n in zip(got, (q2, k2, v2, e2), "qkve"):
        close(a, b.grad, f"d{n}", tol)


@pytest.mark.parametrize("C,groups", [(512, 1), (1024, 1), (100, 1), (32, 16)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layer_norm_backward_kernel(C, groups, dt):
    from anemoi_core_b200 import autograd as AG

    g = torch.Generator().manual_seed(C + groups)
    M = 777
    x = (torch.randn(M, groups * C, generator=g) * 2 + 0.5).to(dt).cuda().requires_grad_()
    wt, b = (torch.randn(C, generator=g).cuda().requires_grad_() for _ in range(2))
    w = torch.randn(M, groups * C, generator=g).cuda()
    y = AG.layer_norm(x, wt, b, 1e-5, dt, groups)
    (y.float() * w).sum().backward()
    x2, w2, b2 = x.detach().float().requires_grad_(), wt.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    ref = torch.nn.functional.layer_norm(x2.view(M, groups, C), (C,), w2, b2, 1e-5).view(M, groups * C)
    (ref * w).sum().backward()
    tol = 1e-4 if dt == torch.float32 else 3e-2
    close(y, ref, "layer_norm forward", tol)
    close(x.grad, x2.grad, "dx", tol)
    close(wt.grad, w2.grad, "dgamma", 5e-4 if dt == torch.float32 else 3e-2)
    close(b.grad, b2.grad, "dbeta", 5e-4 if dt == torch.float32 else 3e-2)


def test_linear_and_gelu_backward():
    from anemoi_core_b200 import autograd as AG

    g = torch.Generator().manual_seed(5)
    for dt, K, N in ((torch.float32, 37, 50), (torch.bfloat16, 512, 256), (torch.bfloat16, 11, 64)):
        x = torch.randn(1000, K, generator=g).cuda().requires_grad_()
        lin = torch.nn.Linear(K, N).cuda()
        w = torch.randn(1000, N, generator=g).cuda()
        y = AG.linear(x, lin.weight, lin.bias, dt, gelu=True)
        (y.float() * w).sum().backward()
        got = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
        x.grad = None
        lin.zero_grad()
        ref = torch.nn.functional.gelu(torch.nn.functional.linear(x, lin.weight, lin.bias))
        (ref * w).sum().backward()
        tol = 1e-4 if dt == torch.float32 else 3e-2
        close(y, ref, f"linear+gelu forward {dt}", tol)
        for a, b, n in zip(got, (x.grad, lin.weight.grad, lin.bias.grad), ("dx", "dW", "db")):
            close(a, b, f"linear {n} {dt}", tol)


def test_training_step_bf16_autocast_cfg2_width():
    """One training step of a 2-layer 512-wide GraphTransformer processor under bf16 autocast: finite gradients for every parameter, close
    to the fp32 gradients of the oracle (rel-L2 <= 5e-2 on the input gradient)."""
    from anemoi_core_b200.layers import GraphTransformerProcessor
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph("o32", 4)
    torch.manual_seed(0)
    m = GraphTransformerProcessor(num_layers=2, num_channels=512, num_chunks=1, num_heads=16, mlp_hidden_ratio=4, edge_dim=gr["edge_dim"]).train()
    sd = {k: v.clone().requires_grad_() for k, v in m.state_dict().items()}
    x = torch.randn(gr["n_mesh"], 512, generator=torch.Generator().manual_seed(1))
    w = torch.randn(gr["n_mesh"], 512, generator=torch.Generator().manual_seed(2))
    xr = x.clone().requires_grad_()
    ref = R.gt_processor(sd, xr, gr["proc_attr"], gr["proc_index"], 2, 16)
    (ref * w).sum().backward()
    m = m.cuda()
    xc = x.cuda().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(xc, 1, shard1(x.shape[0]), gr["proc_attr"].cuda(), gr["proc_index"].cuda())
    (y.float() * w.cuda()).sum().backward()
    l2 = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()  # noqa: E731
    assert l2(y, ref.detach()) <= 2e-2
    assert l2(xc.grad, xr.grad) <= 5e-2, l2(xc.grad, xr.grad)
    big = max(t.grad.norm().item() for t in sd.values())
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if sd[n].grad.norm().item() > 1e-3 * big:  # (gradients that vanish by construction, e.g. lin_key.bias, only hold rounding noise)
            assert l2(p.grad, sd[n].grad) <= 1e-1, (n, l2(p.grad, sd[n].grad))


# ---- round 2: gated feed-forward variants and ConditionalLayerNorm in training (tests/golden/grads_r2.pt, oracle/gen_grad_golden_r2.py) ----
@pytest.mark.parametrize("name", ["gt_processor_glu", "gt_processor_swiglu", "gt_processor_geglu", "gt_processor_reglu", "gnn_processor_swiglu",
                                  "gnn_processor_geglu"])  # fmt: skip
def test_gated_mlp_gradients_match_reference(golden, name):
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    c = golden("grads_r2")[name]
    m = (GNNProcessor if name.startswith("gnn") else GraphTransformerProcessor)(**c["cfg"])
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    x, ea = c["x"].cuda().requires_grad_(), c["edge_attr"].cuda().requires_grad_()
    y = m(x, 1, shard1(x.shape[0]), ea, c["edge_index"].cuda())
    close(y, c["y"], f"{name} forward")
    (y * c["w"].cuda()).sum().backward()
    close(x.grad, c["grads"]["x"], f"{name} dx")
    close(ea.grad, c["grads"]["edge_attr"], f"{name} dedge_attr")
    check_grads(m, c["grads"], name)


@pytest.mark.parametrize("act", ["glu", "swiglu", "geglu", "reglu"])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_glu_combine_backward_kernel(act, dt):
    from anemoi_core_b200 import autograd as AG

    g = torch.Generator().manual_seed(5)
    M, H = 333, 96
    gv = torch.randn(M, 2 * H, generator=g).to(dt)
    w = torch.randn(M, H, generator=g).to(dt)
    fn = {"glu": torch.sigmoid, "swiglu": torch.nn.functional.silu, "geglu": torch.nn.functional.gelu, "reglu": torch.relu}[act]
    ref_in = gv.float().clone().requires_grad_()
    (fn(ref_in[:, :H]) * ref_in[:, H:] * w.float()).sum().backward()
    x = gv.detach().clone().cuda().requires_grad_()
    (AG.glu_combine(x, act).float() * w.cuda().float()).sum().backward()
    tol = 2e-5 if dt == torch.float32 else 2**-6
    assert (x.grad.float().cpu() - ref_in.grad).abs().max().item() <= tol * ref_in.grad.abs().max().item()


def _cond_kernels(dc):
    return {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": dc, "zero_init": False}}


def test_conditional_layer_norm_processor_gradients_match_reference(golden):
    from anemoi_core_b200.layers import GraphTransformerProcessor

    c = golden("grads_r2")["gt_processor_condln"]
    m = GraphTransformerProcessor(layer_kernels=_cond_kernels(c["condition_shape"]), **c["cfg"])
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    x, ea, cond = (c[k].cuda().requires_grad_() for k in ("x", "edge_attr", "cond"))
    y = m(x, 1, shard1(), ea, c["edge_index"].cuda(), cond=cond)
    close(y, c["y"], "condln forward")
    (y * c["w"].cuda()).sum().backward()
    for k, t in (("x", x), ("edge_attr", ea), ("cond", cond)):
        close(t.grad, c["grads"][k], f"condln d{k}")
    check_grads(m, c["grads"], "condln processor")


def test_conditional_layer_norm_mapper_gradients_match_reference(golden):
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.layers import GraphTransformerForwardMapper

    c = golden("grads_r2")["gt_forward_mapper_condln"]
    m = GraphTransformerForwardMapper(layer_kernels=_cond_kernels(c["condition_shape"]), **c["cfg"])
    m.load_state_dict(c["sd"], strict=True)
    m = m.cuda().train()
    xs, xd, ea, cs, cd = (c[k].cuda().requires_grad_() for k in ("x_src", "x_dst", "edge_attr", "cond_src", "cond_dst"))
    out = m((xs, xd), 1, BipartiteGraphShardInfo(), ea, c["edge_index"].cuda(), cond=(cs, cd))
    ys = [out[1], out[0]]
    for y, yr in zip(ys, c["y"]):
        close(y, yr, "condln mapper forward")
    sum((y * w.cuda()).sum() for y, w in zip(ys, c["w"])).backward()
    for k, t in (("x_src", xs), ("x_dst", xd), ("edge_attr", ea), ("cond_src", cs), ("cond_dst", cd)):
        if c["grads"][k] is None:
            assert t.grad is None or t.grad.abs().max().item() == 0.0
        else:
            close(t.grad, c["grads"][k], f"condln mapper d{k}")
    check_grads(m, c["grads"], "condln mapper")


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(1, 8), (10, 512), (33, 7), (5000, 100), (40962, 512), (70001, 2560), (300000, 64)])
def test_col_sum_kernel(shape, dt):
    """ops.col_sum (the bias gradient: column sums of a cotangent, fp32, two-stage, no atomics) against a float64 sum; a strided view (the
    pitch of a wider matrix) and run-to-run bit-identity."""
    from anemoi_core_b200 import ops

    M, N = shape
    torch.manual_seed(M + N)
    wide = torch.randn(M, (N + 24 + 7) // 8 * 8, device="cuda").to(dt)  # rows of whole 16-byte groups; ragged N inside them
    views = [wide[:, :N], wide[:, 16 : 16 + N]]
    if (N * wide.element_size()) % 16 == 0:
        views.append(wide[:, :N].contiguous())
    for x in views:
        got = ops.col_sum(x)
        assert got.dtype == torch.float32 and got.shape == (N,)
        ref = x.double().sum(0)
        bound = 4e-6 * x.double().abs().sum(0) + 1e-6  # fp32 accumulation of M terms in a fixed tree
        assert ((got.double() - ref).abs() <= bound).all(), f"{shape} {dt}: {(got.double() - ref).abs().max().item():.3e}"
        assert torch.equal(got, ops.col_sum(x))
    assert torch.equal(ops.col_sum(torch.empty(0, 16, device="cuda", dtype=dt)), torch.zeros(16, device="cuda"))


def test_bias_gradient_through_col_sum(monkeypatch):
    """LinearFn's db with ANEMOI_B200_COL_SUM on == PyTorch's sum(0) path."""
    from anemoi_core_b200 import autograd as AG

    torch.manual_seed(2)
    x = torch.randn(3000, 256, device="cuda")
    w = torch.randn(512, 256, device="cuda", requires_grad=True)
    b = torch.randn(512, device="cuda", requires_grad=True)
    g = torch.randn(3000, 512, device="cuda")
    out = {}
    for on in (False, True):
        monkeypatch.setattr(AG, "COL_SUM", on)
        for dt in (torch.float32, torch.bfloat16):
            w.grad = b.grad = None
            (AG.LinearFn.apply(x, w, b, True, dt).float() * g).sum().backward()
            out[on, dt] = b.grad.clone()
    close(out[True, torch.float32], out[False, torch.float32], "db fp32", 1e-5)
    close(out[True, torch.bfloat16], out[False, torch.bfloat16], "db bf16", 1e-5)


def test_gelu_backward_two_mufu_form(monkeypatch):
    """ANEMOI_B200_GELU_BWD_FAST: gelu'(x) through two exp2 (bf16 cotangents only) against the derivative of torch's exact-erf GELU; the bound is
    1.3e-5 on gelu' plus the bf16 rounding of the result."""
    from anemoi_core_b200 import ops

    torch.manual_seed(4)
    x = (torch.randn(4099, 512, device="cuda") * 3).to(torch.bfloat16)
    x[0, :8] = torch.tensor([0.0, -0.0, 1e-4, -1e-4, 12.0, -12.0, 40.0, -40.0], device="cuda").to(torch.bfloat16)
    dy = torch.randn(4099, 512, device="cuda").to(torch.bfloat16)
    xf = x.double().requires_grad_()
    torch.nn.functional.gelu(xf).backward(dy.double())
    monkeypatch.setattr(ops, "GELU_BWD_FAST", True)
    got = ops.gelu(x, dy)
    err = (got.double() - xf.grad).abs()
    assert (err <= 2e-5 * dy.double().abs() + 2.0**-8 * xf.grad.abs() + 1e-30).all(), err.max().item()
    x32, dy32 = x.float(), dy.float()  # fp32 keeps the exact form, switch or not
    monkeypatch.setattr(ops, "GELU_BWD_FAST", False)
    ref32 = ops.gelu(x32, dy32)
    monkeypatch.setattr(ops, "GELU_BWD_FAST", True)
    assert torch.equal(ops.gelu(x32, dy32), ref32)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_linear_residual_in_epilogue(dt):
    """LinearFn with the residual added in the GEMM epilogue (ANEMOI_B200_TRAIN_FUSE_RES) == the Linear followed by a PyTorch add: forward and all
    four gradients (the residual's is the cotangent itself)."""
    from anemoi_core_b200 import autograd as AG

    torch.manual_seed(8)
    g = torch.randn(3001, 512, device="cuda")
    res = {}
    for fused in (False, True):
        torch.manual_seed(9)
        x = torch.randn(3001, 256, device="cuda", requires_grad=True)
        w = torch.randn(512, 256, device="cuda", requires_grad=True)
        b = torch.randn(512, device="cuda", requires_grad=True)
        r = torch.randn(3001, 512, device="cuda").to(dt).requires_grad_()
        y = AG.linear(x, w, b, dt, False, r) if fused else AG.linear(x, w, b, dt) + r
        (y.float() * g).sum().backward()
        res[fused] = (y, x.grad, w.grad, b.grad, r.grad)
    tol = 1e-5 if dt == torch.float32 else 1e-2  # bf16: one rounding of (acc + bias + residual) against two
    for a, c, n in zip(res[True], res[False], ("y", "dx", "dW", "db", "dres")):
        close(a, c, f"fused residual {n} {dt}", tol if n == "y" else 1e-6)
