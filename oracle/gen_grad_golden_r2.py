"""oracle/gen_grad_golden_r2.py — more gradient fixtures from the UNMODIFIED reference modules (PyTorch autograd on CPU, fp32): the gated
feed-forward variants (layers/mlp.py:38-94) and ConditionalLayerNorm kernels (layers/normalization.py:34-94) in training.

*** TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference). ***     python oracle/gen_grad_golden_r2.py

Same layout as tests/golden/grads.pt (oracle/gen_grad_golden.py): seeded inputs, reference ``state_dict``, cotangent ``w``, forward output,
gradients of every parameter, of the node inputs, the edge attributes and (where there is one) the conditioning tensor.
-> tests/golden/grads_r2.pt
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

import gen_grad_golden as G  # noqa: E402  (sets up the import path of the reference + stand-ins)
import torch  # noqa: E402
from anemoi.utils.config import DotDict  # noqa: E402

from anemoi.models.distributed.shapes import BipartiteGraphShardInfo  # noqa: E402
from anemoi.models.distributed.shapes import GraphShardInfo  # noqa: E402
from anemoi.models.layers.mapper import GraphTransformerForwardMapper  # noqa: E402
from anemoi.models.layers.processor import GNNProcessor  # noqa: E402
from anemoi.models.layers.processor import GraphTransformerProcessor  # noqa: E402


def case_gated(kind, impl, seed):
    n, e, edge_dim = 80, 300, 5
    torch.manual_seed(seed)
    if kind == "gt":
        cfg = dict(num_layers=1, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=4, edge_dim=edge_dim, mlp_implementation=impl)
        m = GraphTransformerProcessor(layer_kernels=None, graph_attention_backend="pyg", **cfg)
    else:
        cfg = dict(num_channels=32, num_layers=1, num_chunks=1, mlp_extra_layers=0, edge_dim=edge_dim, mlp_implementation=impl)
        m = GNNProcessor(layer_kernels=None, **cfg)
    m = G.randomise(m, seed).train()
    ei, ea = G.rand_graph(n, n, e, edge_dim, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(n, cfg["num_channels"], generator=g).requires_grad_()
    ea = ea.requires_grad_()
    y = m(x, 1, GraphShardInfo(nodes=[n], edges=None), ea, ei, None)
    w = torch.randn(y.shape, generator=g)
    return {"cfg": cfg, "sd": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(), "edge_attr": ea.detach().clone(),
            "edge_index": ei, "w": w, "y": y.detach().clone(), "grads": G.grads_of(m, [y], [w], {"x": x, "edge_attr": ea})}  # fmt: skip


def case_condln(seed=71):
    n, e, c, heads, layers, edge_dim, dc = 70, 180, 64, 4, 2, 5, 16
    lk = DotDict({"LayerNorm": {"_target_": "anemoi.models.layers.normalization.ConditionalLayerNorm", "condition_shape": dc, "zero_init": False}})
    torch.manual_seed(seed)
    cfg = dict(num_layers=layers, num_channels=c, num_chunks=1, num_heads=heads, mlp_hidden_ratio=4, edge_dim=edge_dim)
    m = G.randomise(GraphTransformerProcessor(layer_kernels=lk, graph_attention_backend="pyg", **cfg), seed).train()
    ei, ea = G.rand_graph(n, n, e, edge_dim, seed)
    g = torch.Generator().manual_seed(seed + 1)
    x, cond = torch.randn(n, c, generator=g).requires_grad_(), torch.randn(n, dc, generator=g).requires_grad_()
    ea = ea.requires_grad_()
    y = m(x, 1, GraphShardInfo(nodes=None, edges=None), ea, ei, None, cond=cond)
    w = torch.randn(y.shape, generator=g)
    proc = {"cfg": cfg, "condition_shape": dc, "sd": {k: v.detach().clone() for k, v in m.state_dict().items()}, "x": x.detach().clone(),
            "cond": cond.detach().clone(), "edge_attr": ea.detach().clone(), "edge_index": ei, "w": w, "y": y.detach().clone(),
            "grads": G.grads_of(m, [y], [w], {"x": x, "edge_attr": ea, "cond": cond})}  # fmt: skip
    # forward mapper with (cond_src, cond_dst) (block.py:978-1023)
    n_src, n_dst, in_src, in_dst = 90, 70, 10, 6
    torch.manual_seed(seed + 2)
    mcfg = dict(in_channels_src=in_src, in_channels_dst=in_dst, hidden_dim=c, num_chunks=1, num_heads=heads, mlp_hidden_ratio=4, edge_dim=edge_dim)
    mm = G.randomise(GraphTransformerForwardMapper(layer_kernels=lk, graph_attention_backend="pyg", **mcfg), seed).train()
    mei, mea = G.rand_graph(n_src, n_dst, 200, edge_dim, seed + 3)
    xs, xd = torch.randn(n_src, in_src, generator=g).requires_grad_(), torch.randn(n_dst, in_dst, generator=g).requires_grad_()
    cs, cd = torch.randn(n_src, dc, generator=g).requires_grad_(), torch.randn(n_dst, dc, generator=g).requires_grad_()
    mea = mea.requires_grad_()
    ys_, yd = mm((xs, xd), 1, BipartiteGraphShardInfo(src_nodes=None, dst_nodes=None, edges=None), mea, mei, None, cond=(cs, cd))
    ws = [torch.randn(yd.shape, generator=g), torch.randn(ys_.shape, generator=g)]
    mapper = {"cfg": mcfg, "condition_shape": dc, "sd": {k: v.detach().clone() for k, v in mm.state_dict().items()}, "x_src": xs.detach().clone(),
              "x_dst": xd.detach().clone(), "cond_src": cs.detach().clone(), "cond_dst": cd.detach().clone(), "edge_attr": mea.detach().clone(),
              "edge_index": mei, "w": ws, "y": [yd.detach().clone(), ys_.detach().clone()],
              "grads": G.grads_of(mm, [yd, ys_], ws, {"x_src": xs, "x_dst": xd, "edge_attr": mea, "cond_src": cs, "cond_dst": cd})}  # fmt: skip
    return proc, mapper


def main():
    out = {"kind": "grads_r2"}
    for i, impl in enumerate(("glu", "swiglu", "geglu", "reglu")):
        out[f"gt_processor_{impl}"] = case_gated("gt", impl, 61 + i)
    out["gnn_processor_swiglu"] = case_gated("gnn", "swiglu", 66)
    out["gnn_processor_geglu"] = case_gated("gnn", "geglu", 67)
    out["gt_processor_condln"], out["gt_forward_mapper_condln"] = case_condln()
    torch.save(out, os.path.join(ROOT, "tests", "golden", "grads_r2.pt"))
    for k, v in out.items():
        if isinstance(v, dict):
            print(k, len(v["grads"]["params"]), "param grads;", {n: (None if g is None else float(g.abs().mean())) for n, g in v["grads"].items() if n != "params"})


if __name__ == "__main__":
    main()
