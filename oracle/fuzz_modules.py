"""oracle/fuzz_modules.py — randomised module-level comparison with the UNMODIFIED reference modules, on CPU.

*** TEST INFRASTRUCTURE (run by tests/test_host_logic.py in a subprocess).  Never imported by anemoi_core_b200. ***

The committed goldens pin the drop-in modules at 23 fixed configurations.  This walks RANDOM configurations of the same constructors —
channels, heads, layers, edge_dim, qk_norm, mlp_implementation, edge_pre_mlp, attn_channels, mlp_extra_layers, (un)sorted edges, bipartite
sizes — builds the reference module (anemoi.models.layers.{processor,mapper}, pyg attention backend, imported from /root/reference/models/src or
baseline/_ref with oracle/standins), loads its ``state_dict`` into ours with ``strict=True`` and compares
  * the inference path (``eval()``, ``no_grad``: LayerNorm folds, packed weights, folded lin_edge, gather-add GEMM, one-kernel GraphConv routing),
  * the training path (``train()``: layers/_train.py over autograd.py): output and the gradients of every parameter, the node inputs and the
    edge attributes,
with ``tests/_cpu_ops.py`` standing in for the CUDA entry points (plain fp32 PyTorch statements of each fused op): what is compared is the HOST
side of the product against the reference itself.  Bar: 1e-4 of each tensor's scale (SURVEY.md §8d fp32 bar; gradients that are zero by
construction are floored at 1e-3 of the largest gradient, like tests/test_gpu_backward.py).
    python oracle/fuzz_modules.py [cases]   ->   one JSON line {"cases": n, "worst": {...}, "failures": [...]}"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle.reference_step import reference_root  # noqa: E402


def _rel(a, b, floor=1e-6):
    return ((a.detach().float() - b.detach().float()).abs().max() / max(b.detach().abs().max().item(), floor)).item()


def _graph(g, n_src, n_dst, n_edges, edge_dim, sort=True):
    ei = torch.stack([torch.randint(0, n_src, (n_edges,), generator=g), torch.randint(0, n_dst, (n_edges,), generator=g)])
    if sort:
        ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
    return ei, torch.randn(n_edges, edge_dim, generator=g)


def _randomise(m, g):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    return m


def _pick(g, options):
    return options[int(torch.randint(0, len(options), (1,), generator=g))]


def _cond_kernels(dc: int):
    """(reference, ours) ``layer_kernels`` selecting ConditionalLayerNorm (normalization.py:34-94) for the LayerNorm kernel, or (None, None)."""
    if not dc:
        return None, None
    from anemoi.utils.config import DotDict

    kw = {"condition_shape": dc, "zero_init": False}
    return (DotDict({"LayerNorm": {"_target_": "anemoi.models.layers.normalization.ConditionalLayerNorm", **kw}}),
            {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", **kw}})  # fmt: skip


def _compare(ref_m, our_m, call_ref, call_ours, inputs, g, tag, worst, fails):
    """``inputs``: dict name -> tensor (differentiable inputs); call_*(module, **inputs) -> tensor or tuple of tensors."""
    our_m.load_state_dict({k: v.detach().clone() for k, v in ref_m.state_dict().items()}, strict=True)

    def outs(y):
        return [t for t in (y if isinstance(y, (tuple, list)) else (y,)) if torch.is_tensor(t) and t.is_floating_point()]

    ref_m.eval(), our_m.eval()
    with torch.no_grad():
        yr, yo = outs(call_ref(ref_m, **inputs)), outs(call_ours(our_m, **inputs))
    errs = {"inference": max(_rel(a, b) for a, b in zip(yo, yr))}
    ref_m.train(), our_m.train()
    res = {}
    for who, m, call in (("ref", ref_m, call_ref), ("ours", our_m, call_ours)):
        m.zero_grad()
        xs = {k: v.detach().clone().requires_grad_() for k, v in inputs.items()}
        ys = outs(call(m, **xs))
        if who == "ref":
            ws = [torch.randn(y.shape, generator=g) for y in ys]
        sum((y * w).sum() for y, w in zip(ys, ws) if y.requires_grad).backward()
        res[who] = (ys, {k: v.grad for k, v in xs.items()}, {k: p.grad for k, p in m.named_parameters()})
    errs["training forward"] = max(_rel(a, b) for a, b in zip(res["ours"][0], res["ref"][0]))
    errs["input gradients"] = max([_rel(res["ours"][1][k], gr) for k, gr in res["ref"][1].items() if gr is not None] or [0.0])
    pg = {k: v for k, v in res["ref"][2].items() if v is not None}
    floor = 1e-3 * max(v.abs().max().item() for v in pg.values())
    missing = [k for k in pg if res["ours"][2].get(k) is None]
    errs["parameter gradients"] = max(_rel(res["ours"][2][k], v, floor) for k, v in pg.items() if k not in missing)
    for what, e in errs.items():
        worst[what] = max(worst.get(what, 0.0), e)
        if not e <= 1e-4 and len(fails) < 10:
            fails.append(f"{tag}: {what} {e:.3e}")
    if missing and len(fails) < 10:
        fails.append(f"{tag}: no gradient for {missing[:3]}")


def main(n_cases: int) -> dict:
    root = reference_root()
    if root is None:
        return {"unavailable": "reference not found (neither /root/reference/models/src nor baseline/_ref)"}
    for p in (root, os.path.join(HERE, "standins")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from anemoi.models.distributed.shapes import BipartiteGraphShardInfo as RBSI
    from anemoi.models.distributed.shapes import GraphShardInfo as RGSI
    from anemoi.models.layers import mapper as RM
    from anemoi.models.layers import processor as RP

    import _cpu_ops

    _cpu_ops.install()
    from anemoi_core_b200 import layers as L
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    torch.set_num_threads(4)
    g = torch.Generator().manual_seed(20260)
    worst, fails, kinds = {}, [], {}
    for case in range(n_cases):
        kind = ("gt_processor", "gnn_processor", "gt_forward_mapper", "gt_backward_mapper", "gnn_forward_mapper", "gnn_backward_mapper")[case % 6]
        kinds[kind] = kinds.get(kind, 0) + 1
        torch.manual_seed(1000 + case)
        edge_dim = _pick(g, [3, 5, 11])
        if kind == "gt_processor":
            heads = _pick(g, [2, 4, 8])
            cfg = dict(num_layers=_pick(g, [1, 2, 3]), num_channels=heads * _pick(g, [8, 16]), num_heads=heads, edge_dim=edge_dim, qk_norm=_pick(g, [False, True]),
                       mlp_implementation=_pick(g, ["mlp", "mlp", "glu", "swiglu", "geglu", "reglu"]))  # fmt: skip
            if _pick(g, [False, True]):
                cfg["edge_pre_mlp"] = True
            if _pick(g, [False, True]):
                cfg["attn_channels"] = 2 * cfg["num_channels"]
            sort = _pick(g, [True, True, False])
            n = int(torch.randint(20, 120, (1,), generator=g))
            ei, ea = _graph(g, n, n, int(torch.randint(n, 6 * n, (1,), generator=g)), edge_dim, sort)
            cfg.update(num_chunks=_pick(g, [c_ for c_ in (1, 2, 3) if cfg["num_layers"] % c_ == 0]), mlp_hidden_ratio=_pick(g, [2, 4]))
            dc = _pick(g, [0, 0, 0, 8])  # one in four: ConditionalLayerNorm kernels driven by a per-node conditioning tensor (cond=)
            lk_ref, lk_ours = _cond_kernels(dc)
            ref_m = _randomise(RP.GraphTransformerProcessor(layer_kernels=lk_ref, graph_attention_backend="pyg", **cfg), g)
            our_m = L.GraphTransformerProcessor(layer_kernels=lk_ours, **cfg)
            inputs = {"x": torch.randn(n, cfg["num_channels"], generator=g), "edge_attr": ea}
            if dc:
                inputs["cond"] = torch.randn(n, dc, generator=g)
                cfg["condition_shape"] = dc
            call_ref = lambda m, x, edge_attr, **kw: m(x, 1, RGSI(nodes=None, edges=None), edge_attr, ei, None, edges_are_dst_sorted=sort, **kw)  # noqa: E731
            call_ours = lambda m, x, edge_attr, **kw: m(x, 1, GraphShardInfo(), edge_attr, ei, None, edges_are_dst_sorted=sort, **kw)  # noqa: E731
        elif kind == "gnn_processor":
            cfg = dict(num_channels=_pick(g, [16, 32, 48, 64]), num_layers=_pick(g, [1, 2, 3]), mlp_extra_layers=_pick(g, [0, 0, 1]), edge_dim=edge_dim,
                       mlp_implementation=_pick(g, ["mlp", "mlp", "swiglu"]))  # fmt: skip
            n = int(torch.randint(20, 120, (1,), generator=g))
            ei, ea = _graph(g, n, n, int(torch.randint(n, 6 * n, (1,), generator=g)), edge_dim)
            cfg["num_chunks"] = _pick(g, [c_ for c_ in (1, 2, 3) if cfg["num_layers"] % c_ == 0])
            ref_m = _randomise(RP.GNNProcessor(layer_kernels=None, **cfg), g)
            our_m = L.GNNProcessor(**cfg)
            inputs = {"x": torch.randn(n, cfg["num_channels"], generator=g), "edge_attr": ea}
            call_ref = lambda m, x, edge_attr: m(x, 1, RGSI(nodes=[n], edges=None), edge_attr, ei, None)  # noqa: E731
            call_ours = lambda m, x, edge_attr: m(x, 1, GraphShardInfo(nodes=[n]), edge_attr, ei, None)  # noqa: E731
        else:
            n_src, n_dst = int(torch.randint(20, 150, (1,), generator=g)), int(torch.randint(20, 150, (1,), generator=g))
            ei, ea = _graph(g, n_src, n_dst, int(torch.randint(n_dst, 5 * n_dst, (1,), generator=g)), edge_dim)
            gt = kind.startswith("gt")
            heads = _pick(g, [2, 4])
            c = heads * _pick(g, [8, 16]) if gt else _pick(g, [16, 32, 48])
            fwd = "forward" in kind
            in_src, in_dst = (int(torch.randint(3, 20, (1,), generator=g)), int(torch.randint(3, 20, (1,), generator=g))) if fwd else (c, int(torch.randint(3, 20, (1,), generator=g)))
            cfg = dict(in_channels_src=in_src, in_channels_dst=in_dst, hidden_dim=c, edge_dim=edge_dim)
            if not fwd:
                cfg["out_channels_dst"] = int(torch.randint(2, 12, (1,), generator=g))
            if gt:
                cfg.update(num_heads=heads, qk_norm=_pick(g, [False, True]))
                cfg.update(mlp_hidden_ratio=_pick(g, [2, 4]), num_chunks=_pick(g, [1, 2, 4]))
                extra_ref, extra_ours = dict(layer_kernels=None, graph_attention_backend="pyg"), {}
            else:
                extra_ref, extra_ours = dict(mlp_extra_layers=0, num_chunks=1, layer_kernels=None), dict(mlp_extra_layers=0, num_chunks=1)
                if not fwd:
                    cfg["in_channels_dst"] = c  # the GNN decoder's destination rows are hidden rows (mapper.py:1045-1054)
            name = {"gt_forward_mapper": "GraphTransformerForwardMapper", "gt_backward_mapper": "GraphTransformerBackwardMapper",
                    "gnn_forward_mapper": "GNNForwardMapper", "gnn_backward_mapper": "GNNBackwardMapper"}[kind]  # fmt: skip
            dc = _pick(g, [0, 0, 0, 8]) if kind == "gt_forward_mapper" else 0
            if dc:
                extra_ref["layer_kernels"], extra_ours["layer_kernels"] = _cond_kernels(dc)
            ref_m = _randomise(getattr(RM, name)(**cfg, **extra_ref), g)
            our_m = getattr(L, name)(**cfg, **extra_ours)
            inputs = {"x_src": torch.randn(n_src, cfg["in_channels_src"], generator=g), "x_dst": torch.randn(n_dst, cfg["in_channels_dst"], generator=g), "edge_attr": ea}
            if dc:
                inputs.update(cond_src=torch.randn(n_src, dc, generator=g), cond_dst=torch.randn(n_dst, dc, generator=g))
                cfg["condition_shape"] = dc

            def _kw(cond_src=None, cond_dst=None):
                return {} if cond_src is None else {"cond": (cond_src, cond_dst)}

            call_ref = lambda m, x_src, x_dst, edge_attr, **c: m((x_src, x_dst), 1, RBSI(src_nodes=None, dst_nodes=None, edges=None), edge_attr, ei, None, **_kw(**c))  # noqa: E731
            call_ours = lambda m, x_src, x_dst, edge_attr, **c: m((x_src, x_dst), 1, BipartiteGraphShardInfo(), edge_attr, ei, None, **_kw(**c))  # noqa: E731
        tag = f"case {case} {kind} {cfg}"
        if "condition_shape" in cfg:
            kinds["with ConditionalLayerNorm"] = kinds.get("with ConditionalLayerNorm", 0) + 1
        try:
            _compare(ref_m, our_m, call_ref, call_ours, inputs, g, tag, worst, fails)
        except Exception as e:  # noqa: BLE001
            import traceback

            if len(fails) < 10:
                fails.append(f"{tag}: {type(e).__name__}: {e} | {traceback.format_exc().splitlines()[-3:]}")
    return {"cases": n_cases, "kinds": kinds, "worst": {k: float(f"{v:.3e}") for k, v in worst.items()}, "failures": fails, "reference": root}


if __name__ == "__main__":
    print(json.dumps(main(int(sys.argv[1]) if len(sys.argv) > 1 else 60)))
