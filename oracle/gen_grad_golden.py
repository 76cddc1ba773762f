"""oracle/gen_grad_golden.py — gradient fixtures from the UNMODIFIED reference modules (PyTorch autograd on CPU, fp32).

*** TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference). ***     python oracle/gen_grad_golden.py

For each case: seeded inputs, the reference ``state_dict``, a seeded cotangent ``w`` (loss = sum(y * w)), the forward output and the
gradients of every parameter, of the node inputs and of the edge attributes.  Cases: GraphTransformerProcessor (with and without
qk_norm), GNNProcessor, the four mappers, and GraphTransformerConv at the reference's Triton-parity shapes
(models/tests/integration/triton/test_triton_gt.py:49-57, 117-184: forward + backward, atol 1e-4).  -> tests/golden/grads.pt
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(1, "/root/reference/models/src")
sys.path.insert(2, ROOT)

import torch  # noqa: E402

from anemoi.models.distributed.shapes import BipartiteGraphShardInfo  # noqa: E402
from anemoi.models.distributed.shapes import GraphShardInfo  # noqa: E402
from anemoi.models.layers.conv import GraphTransformerConv  # noqa: E402
from anemoi.models.layers.mapper import GNNBackwardMapper  # noqa: E402
from anemoi.models.layers.mapper import GNNForwardMapper  # noqa: E402
from anemoi.models.layers.mapper import GraphTransformerBackwardMapper  # noqa: E402
from anemoi.models.layers.mapper import GraphTransformerForwardMapper  # noqa: E402
from anemoi.models.layers.processor import GNNProcessor  # noqa: E402
from anemoi.models.layers.processor import GraphTransformerProcessor  # noqa: E402


def rand_graph(n_src, n_dst, n_edges, edge_dim, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n_src, (n_edges,), generator=g), torch.randint(0, n_dst, (n_edges,), generator=g)])
    ei = ei[:, torch.sort(ei[1], stable=True)[1]]
    return ei, torch.randn(n_edges, edge_dim, generator=g)


def randomise(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return m


def grads_of(m, y, w, extra):
    loss = sum((yi * wi).sum() for yi, wi in zip(y, w))
    names = [n for n, _ in m.named_parameters()]
    params = [p for _, p in m.named_parameters()]
    gs = torch.autograd.grad(loss, params + list(extra.values()), allow_unused=True)
    out = {"params": {n: g.clone() for n, g in zip(names, gs[: len(params)]) if g is not None}}
    for (k, _), g in zip(extra.items(), gs[len(params) :]):
        out[k] = None if g is None else g.clone()
    return out


def case_processor(kind, seed, **kw):
    n, e, edge_dim = 100, 400, 5
    torch.manual_seed(seed)
    if kind == "gt":
        cfg = dict(num_layers=2, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=4, edge_dim=edge_dim, qk_norm=kw.get("qk_norm", False))
        m = GraphTransformerProcessor(layer_kernels=None, graph_attention_backend="pyg", **cfg)
    else:
        cfg = dict(num_channels=32, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=edge_dim)
        m = GNNProcessor(layer_kernels=None, **cfg)
    m = randomise(m, seed).train()
    ei, ea = rand_graph(n, n, e, edge_dim, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, cfg["num_channels"], generator=g).requires_grad_()
    ea = ea.requires_grad_()
    y = m(x, 1, GraphShardInfo(nodes=[n], edges=None), ea, ei, None)
    w = torch.randn(y.shape, generator=g)
    return {"cfg": cfg, "sd": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(), "edge_attr": ea.detach().clone(),
            "edge_index": ei, "w": w, "y": y.detach().clone(), "grads": grads_of(m, [y], [w], {"x": x, "edge_attr": ea})}  # fmt: skip


def case_mapper(cls, seed, forward):
    n_src, n_dst, e, edge_dim, c = 90, 70, 260, 3, 64
    in_src, in_dst, out_dst = 10, 6, 5
    torch.manual_seed(seed)
    gt = cls.__name__.startswith("GraphTransformer")
    common = dict(num_heads=4, mlp_hidden_ratio=4, edge_dim=edge_dim, graph_attention_backend="pyg") if gt else dict(mlp_extra_layers=0, edge_dim=edge_dim)
    if forward:
        cfg = dict(in_channels_src=in_src, in_channels_dst=in_dst, hidden_dim=c, num_chunks=1, **common)
    else:
        cfg = dict(in_channels_src=c, in_channels_dst=in_src if gt else c, hidden_dim=c, out_channels_dst=out_dst, num_chunks=1, **common)
    m = randomise(cls(layer_kernels=None, **cfg), seed).train()
    ei, ea = rand_graph(n_src, n_dst, e, edge_dim, seed)
    g = torch.Generator().manual_seed(seed + 1)
    xs = torch.randn(n_src, cfg["in_channels_src"], generator=g).requires_grad_()
    xd = torch.randn(n_dst, cfg["in_channels_dst"], generator=g).requires_grad_()
    ea = ea.requires_grad_()
    out = m((xs, xd), 1, BipartiteGraphShardInfo(src_nodes=None, dst_nodes=None, edges=None), ea, ei, None)
    ys = [out[1]] + ([out[0]] if (forward and not gt) else []) if forward else [out]  # GNN forward mapper also returns the updated src
    ws = [torch.randn(y.shape, generator=g) for y in ys]
    cfg.pop("graph_attention_backend", None)
    return {"cfg": cfg, "sd": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x_src": xs.detach().clone(), "x_dst": xd.detach().clone(),
            "edge_attr": ea.detach().clone(), "edge_index": ei, "w": ws, "y": [y.detach().clone() for y in ys],
            "grads": grads_of(m, ys, ws, {"x_src": xs, "x_dst": xd, "edge_attr": ea})}  # fmt: skip


def case_conv():
    cases = []
    for i, (n_src, n_dst, h, d) in enumerate([(4, 10, 2, 4), (4, 10, 6, 4), (4, 10, 2, 6), (4, 10, 6, 6), (50, 40, 4, 32)]):
        g = torch.Generator().manual_seed(300 + i)
        n_edges = 3 * n_dst
        ei = torch.stack([torch.randint(0, n_src, (n_edges,), generator=g), torch.randint(0, max(n_dst - 2, 1), (n_edges,), generator=g)])
        ei = ei[:, torch.sort(ei[1], stable=True)[1]]
        q, k, v = (torch.randn(n, h, d, generator=g).requires_grad_() for n in (n_dst, n_src, n_src))
        e = torch.randn(n_edges, h, d, generator=g).requires_grad_()
        out = GraphTransformerConv(out_channels=d)(q, k, v, e, ei, size=(n_src, n_dst))
        w = torch.randn(out.shape, generator=g)
        gq, gk, gv, ge = torch.autograd.grad((out * w).sum(), [q, k, v, e])
        cases.append({"q": q.detach(), "k": k.detach(), "v": v.detach(), "e": e.detach(), "edge_index": ei, "w": w, "out": out.detach(), "dq": gq, "dk": gk,
                      "dv": gv, "de": ge})  # fmt: skip
    return cases


def main():
    out = {"kind": "grads",
           "gt_processor": case_processor("gt", 51), "gt_processor_qknorm": case_processor("gt", 52, qk_norm=True), "gnn_processor": case_processor("gnn", 53),
           "gt_forward_mapper": case_mapper(GraphTransformerForwardMapper, 54, True), "gt_backward_mapper": case_mapper(GraphTransformerBackwardMapper, 55, False),
           "gnn_forward_mapper": case_mapper(GNNForwardMapper, 56, True), "gnn_backward_mapper": case_mapper(GNNBackwardMapper, 57, False),
           "gt_conv": case_conv()}  # fmt: skip
    torch.save(out, os.path.join(ROOT, "tests", "golden", "grads.pt"))
    for k, v in out.items():
        if isinstance(v, dict):
            print(k, len(v["grads"]["params"]), "param grads;", {n: (None if g is None else float(g.abs().mean())) for n, g in v["grads"].items() if n != "params"})


if __name__ == "__main__":
    main()
