"""TRAINING on a model-parallel group under Gloo at world size 2, on CPU, with ``tests/_cpu_ops.py`` standing in for the CUDA entry points
(forward AND backward: the stand-ins differentiate their own forward with PyTorch autograd).  What is exercised is the HOST side of the
differentiable sharded path — the autograd halves of the exchanges (``HaloExchangeFn``, ``GatherRowsFn``, ``AllToAllFn``; reference
distributed/graph.py:227-500), the per-rank partial parameter gradients and the input-gradient slices — by comparing one sharded step with
the single-rank step of the SAME stand-in arithmetic: GraphTransformer processor with the "edges" and the "heads" strategy (incl. qk_norm),
GNN processor (all-gathered sources, the gather terms fused into the first edge GEMM, their gradient as segment sums).
The kernels themselves are covered by ``-m gpu`` (tests/test_gpu_backward.py) and the NCCL run by tests/test_gpu_multi.py."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, ret, fn_name="_check"):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        import _cpu_ops

        _cpu_ops.install()
        torch.set_num_threads(2)
        ret[rank] = globals()[fn_name](rank, world)
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


def _graph(n, e, d, seed):
    g = torch.Generator().manual_seed(seed)
    dst = torch.cat([torch.arange(n), torch.randint(0, n, (e - n,), generator=g)])
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), dst])
    ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
    return ei, torch.randn(e, d, generator=g)


def _check(rank, world):
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    n, e, d = 61, 400, 5  # 61 nodes: uneven shards
    ei, ea = _graph(n, e, d, seed=3)
    sizes = get_balanced_partition_sizes(n, world)
    group = dist.group.WORLD
    msgs = []
    for kind in ("gt_edges", "gt_edges_qknorm", "gt_heads", "gnn"):
        torch.manual_seed(0)
        if kind.startswith("gt"):
            c = 32
            m = GraphTransformerProcessor(num_layers=2, num_channels=c, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d,
                                          qk_norm=kind.endswith("qknorm"), shard_strategy="heads" if kind == "gt_heads" else "edges")  # fmt: skip
        else:
            c = 16
            m = GNNProcessor(num_channels=c, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d)
        m.train()
        x0 = torch.randn(n, c, generator=torch.Generator().manual_seed(4))
        w = torch.randn(n, c, generator=torch.Generator().manual_seed(5))
        # single rank
        xf = x0.clone().requires_grad_()
        (m(xf, 1, GraphShardInfo(nodes=[n]), ea, ei) * w).sum().backward()
        ref_p = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        ref_x = xf.grad.clone()
        m.zero_grad()
        # sharded: this rank's rows; parameter gradients are per-rank partial sums (summed over the group, as the trainer does)
        xs = shard_rows(x0, sizes, group).contiguous().clone().requires_grad_()
        y = m(xs, 1, GraphShardInfo(nodes=sizes), ea, ei, group)
        (y * shard_rows(w, sizes, group)).sum().backward()
        rx = shard_rows(ref_x, sizes, group)
        err_x = ((xs.grad - rx).abs().max() / rx.abs().max()).item()
        worst, big = 0.0, max(g.abs().max().item() for g in ref_p.values())
        for k, p in m.named_parameters():
            if k not in ref_p:
                continue
            gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gp)
            worst = max(worst, ((gp - ref_p[k]).abs().max() / max(ref_p[k].abs().max().item(), 1e-3 * big)).item())
        msgs.append((kind, err_x, worst))
    return msgs


def _check_degenerate(rank, world):
    """Training on graphs where a rank's rows receive no edge, every source lives on the last rank, or destinations have no edge at all: the
    autograd halves of the exchanges with empty send / receive lists; sharded gradients == single-rank gradients."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    n, d = 4 * world + 1, 4
    sizes = get_balanced_partition_sizes(n, world)
    group = dist.group.WORLD
    g = torch.Generator().manual_seed(17)

    def graph(src, dst):
        ei = torch.stack([src, dst])
        ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
        return ei, torch.randn(ei.shape[1], d, generator=g)

    graphs = {"sparse": graph(torch.randint(0, n, (n,), generator=g), torch.randint(0, n, (n,), generator=g)),
              "into_first_rank": graph(torch.randint(0, n, (3 * n,), generator=g), torch.randint(0, sizes[0], (3 * n,), generator=g)),
              "from_last_rank": graph(torch.randint(n - sizes[-1], n, (3 * n,), generator=g), torch.randint(0, n, (3 * n,), generator=g))}  # fmt: skip
    msgs = []
    for what, (ei, ea) in graphs.items():
        for kind in ("gt", "gnn"):
            torch.manual_seed(0)
            if kind == "gt":
                c, m = 32, GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d)
            else:
                c, m = 16, GNNProcessor(num_channels=16, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d)
            m.train()
            x0 = torch.randn(n, c, generator=torch.Generator().manual_seed(4))
            w = torch.randn(n, c, generator=torch.Generator().manual_seed(5))
            xf = x0.clone().requires_grad_()
            (m(xf, 1, GraphShardInfo(nodes=[n]), ea, ei) * w).sum().backward()
            ref_p = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
            ref_x = xf.grad.clone()
            m.zero_grad()
            xs = shard_rows(x0, sizes, group).contiguous().clone().requires_grad_()
            (m(xs, 1, GraphShardInfo(nodes=sizes), ea, ei, group) * shard_rows(w, sizes, group)).sum().backward()
            rx = shard_rows(ref_x, sizes, group)
            err_x = ((xs.grad - rx).abs().max() / ref_x.abs().max()).item()
            worst, big = 0.0, max(gr.abs().max().item() for gr in ref_p.values())
            for k, p in m.named_parameters():
                if k not in ref_p:
                    continue
                gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
                dist.all_reduce(gp)
                worst = max(worst, ((gp - ref_p[k]).abs().max() / max(ref_p[k].abs().max().item(), 1e-3 * big)).item())
            msgs.append((f"degenerate_{what}_{kind}", err_x, worst))
    return msgs


def _check_random_configs(rank, world):
    """Random constructor configurations through a SHARDED training step (gated MLPs, qk_norm, edge_pre_mlp, attn_channels, both strategies of the
    GraphTransformer processor, GNN with extra layers / swiglu, several layers and chunks): gradients == the single-rank step."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    group = dist.group.WORLD
    g = torch.Generator().manual_seed(55)

    def pick(options):
        return options[int(torch.randint(0, len(options), (1,), generator=g))]

    msgs = []
    for case in range(8):
        n = int(torch.randint(3 * world, 70, (1,), generator=g))
        d = pick([3, 5])
        ei, ea = _graph(n, int(torch.randint(2 * n, 5 * n, (1,), generator=g)), d, seed=400 + case)
        sizes = get_balanced_partition_sizes(n, world)
        torch.manual_seed(500 + case)
        if case % 2 == 0:
            heads = pick([2, 4])
            c, layers = heads * pick([8, 16]), pick([1, 2])
            cfg = dict(num_layers=layers, num_channels=c, num_chunks=1, num_heads=heads, mlp_hidden_ratio=2, edge_dim=d, qk_norm=pick([False, True]),
                       mlp_implementation=pick(["mlp", "glu", "swiglu", "geglu", "reglu"]), shard_strategy=pick(["edges", "heads"]))  # fmt: skip
            if pick([False, True]):
                cfg["edge_pre_mlp"] = True
            if pick([False, True]):
                cfg["attn_channels"] = 2 * c
            m = GraphTransformerProcessor(**cfg)
        else:
            c, layers = pick([16, 48]), pick([1, 2])
            cfg = dict(num_channels=c, num_layers=layers, num_chunks=1, mlp_extra_layers=pick([0, 1]), edge_dim=d, mlp_implementation=pick(["mlp", "swiglu"]))
            m = GNNProcessor(**cfg)
        m.train()
        x0, w = torch.randn(n, c, generator=g), torch.randn(n, c, generator=g)
        xf = x0.clone().requires_grad_()
        (m(xf, 1, GraphShardInfo(nodes=[n]), ea, ei) * w).sum().backward()
        ref_p = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        ref_x = xf.grad.clone()
        m.zero_grad()
        xs = shard_rows(x0, sizes, group).contiguous().clone().requires_grad_()
        (m(xs, 1, GraphShardInfo(nodes=sizes), ea, ei, group) * shard_rows(w, sizes, group)).sum().backward()
        err_x = ((xs.grad - shard_rows(ref_x, sizes, group)).abs().max() / ref_x.abs().max()).item()
        worst, big = 0.0, max(gr.abs().max().item() for gr in ref_p.values())
        for k, p in m.named_parameters():
            if k not in ref_p:
                continue
            gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gp)
            worst = max(worst, ((gp - ref_p[k]).abs().max() / max(ref_p[k].abs().max().item(), 1e-3 * big)).item())
        msgs.append((f"random_{case}_{type(m).__name__}_{cfg}", err_x, worst))
    return msgs


def _check_model(rank, world):
    """The WHOLE encoder -> processor -> decoder step in training mode, every stage sharded (the call sequence of tests/test_gpu_multi.py's
    training section, here under Gloo on CPU): sharded input / output rows for the GraphTransformer model, replicated in / gathered out for the
    GNN model; gradients of the inputs and of every parameter against the single-rank step."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.model import EncProcDec
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph("o32", mesh_level=3)
    n = gr["n_mesh"]
    group = dist.group.WORLD
    sizes, gsz = get_balanced_partition_sizes(n, world), get_balanced_partition_sizes(gr["n_grid"], world)
    g0, m0 = sum(gsz[:rank]), sum(sizes[:rank])
    gen = torch.Generator().manual_seed(9)
    xg, xm = torch.randn(gr["n_grid"], 20, generator=gen), torch.randn(n, 12, generator=gen)
    wgt = torch.randn(gr["n_grid"], 9, generator=torch.Generator().manual_seed(21))
    msgs = []
    for kind in ("graphtransformer", "gnn"):
        torch.manual_seed(6)
        mt = EncProcDec(kind, in_grid=20, in_mesh=12, out_grid=9, num_channels=32, num_layers=2, edge_dim=gr["edge_dim"], num_heads=4).train()
        xg_f = xg.clone().requires_grad_()
        (mt(xg_f, xm, gr) * wgt).sum().backward()
        ref_p = {k: p.grad.clone() for k, p in mt.named_parameters() if p.grad is not None}
        ref_x = xg_f.grad.clone()
        mt.zero_grad()
        if kind == "graphtransformer":
            xg_s = xg[g0 : g0 + gsz[rank]].clone().requires_grad_()
            y_l = mt(xg_s, xm[m0 : m0 + sizes[rank]].contiguous(), gr, model_comm_group=group, mesh_shards=sizes, grid_shards=gsz,
                     keep_output_sharded=True, inputs_sharded=True)  # fmt: skip
            (y_l * wgt[g0 : g0 + gsz[rank]]).sum().backward()
            gx, rx, scale = xg_s.grad, ref_x[g0 : g0 + gsz[rank]], 1.0
        else:
            xg_s = xg.clone().requires_grad_()
            y_all = mt(xg_s, xm, gr, group, sizes, gsz)  # replicated in / gathered out: every rank back-propagates the same (whole) loss
            (y_all * wgt).sum().backward()
            gx = xg_s.grad.clone()
            dist.all_reduce(gx)
            gx, rx, scale = gx / world, ref_x, 1.0 / world
        err_x = ((gx - rx).abs().max() / rx.abs().max()).item()
        worst, big = 0.0, max(g.abs().max().item() for g in ref_p.values())
        for k, p in mt.named_parameters():
            if k not in ref_p:
                continue
            gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gp)
            worst = max(worst, ((gp * scale - ref_p[k]).abs().max() / max(ref_p[k].abs().max().item(), 1e-3 * big)).item())
        msgs.append((f"encprocdec_{kind}", err_x, worst))
    return msgs


def _check_checkpoint(rank, world):
    """Activation checkpointing of the processors' training path (ANEMOI_B200_ACT_CHECKPOINT; reference layers/processor.py:129-147): a sharded
    training step with every chunk of layers re-run in the backward gives the gradients of the plain step (the exchanges inside a chunk are
    re-run by both ranks alike), and the chunks really went through torch.utils.checkpoint."""
    import torch.utils.checkpoint as tuc

    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor
    from anemoi_core_b200.layers import _train as T

    n, e, d = 61, 400, 5
    ei, ea0 = _graph(n, e, d, seed=3)
    sizes = get_balanced_partition_sizes(n, world)
    group = dist.group.WORLD
    calls = []
    real = tuc.checkpoint
    tuc.checkpoint = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    msgs = []
    for kind in ("gt_edges", "gnn"):
        torch.manual_seed(0)
        if kind == "gt_edges":
            c = 32
            m = GraphTransformerProcessor(num_layers=4, num_channels=c, num_chunks=2, num_heads=4, mlp_hidden_ratio=2, edge_dim=d)
        else:
            c = 16
            m = GNNProcessor(num_channels=c, num_layers=4, num_chunks=2, mlp_extra_layers=0, edge_dim=d)
        m.train()
        x0 = torch.randn(n, c, generator=torch.Generator().manual_seed(4))
        w = shard_rows(torch.randn(n, c, generator=torch.Generator().manual_seed(5)), sizes, group)
        got = {}
        for on in (False, True):
            T.ACT_CHECKPOINT = on
            del calls[:]
            m.zero_grad()
            xs = shard_rows(x0, sizes, group).contiguous().clone().requires_grad_()
            ea = ea0.clone().requires_grad_()
            (m(xs, 1, GraphShardInfo(nodes=sizes), ea, ei, group) * w).sum().backward()
            got[on] = (xs.grad.clone(), ea.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
            assert len(calls) == (2 if on else 0), f"{kind}: {len(calls)} checkpointed chunks with ACT_CHECKPOINT={on}"
        T.ACT_CHECKPOINT = False
        assert set(got[True][2]) == set(got[False][2])
        err_x = max(((got[True][i] - got[False][i]).abs().max() / got[False][i].abs().max()).item() for i in (0, 1))
        err_p = max(((got[True][2][k] - g).abs().max() / max(g.abs().max().item(), 1e-12)).item() for k, g in got[False][2].items())
        msgs.append((f"checkpoint_{kind}", err_x, err_p))
    tuc.checkpoint = real
    return msgs


def _check_model_fixture(rank, world):
    """``AnemoiModelEncProcDec`` in TRAINING mode (differentiable glue: the PyTorch statements of the two assembly kernels, latent sum / skip,
    SkipConnection residual, ReluBounding; graph providers and node attributes with their trainable tensors; the mappers' and the processor's
    differentiable paths) against the gradient fixture of the UNMODIFIED reference model (oracle/gen_grad_golden_model.py ->
    tests/golden/grads_model.pt): output, input gradient, every parameter gradient."""
    from test_model_glue import build_model

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fx = torch.load(os.path.join(gdir, "model_forward.pt"), weights_only=False)
    gr = torch.load(os.path.join(gdir, "grads_model.pt"), weights_only=False)["cases"]
    msgs = []
    for kind in ("graphtransformer", "gnn"):
        m = build_model(fx, kind)
        m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
        m.train()
        x = fx["x"].clone().requires_grad_()
        y = m({"data": x})["data"]
        (y * gr[kind]["w"]).sum().backward()
        ref = gr[kind]["grads"]
        got = {k: p.grad for k, p in m.named_parameters()}
        assert set(ref) <= {k for k, g in got.items() if g is not None}, sorted(set(ref) - {k for k, g in got.items() if g is not None})[:4]
        big = max(g.abs().max().item() for g in ref.values())
        err_p = max(((got[k] - g).abs().max() / max(g.abs().max().item(), 1e-3 * big)).item() for k, g in ref.items())
        err_x = max(((x.grad - gr[kind]["x_grad"]).abs().max() / gr[kind]["x_grad"].abs().max()).item(),
                    ((y - gr[kind]["y"]).abs().max() / gr[kind]["y"].abs().max()).item())  # fmt: skip
        msgs.append((f"model_fixture_{kind}", err_x, err_p))
        # the differentiable glue == the inference glue (pinned to the reference golden) also with an ensemble dimension of 2
        xe = torch.cat([fx["x"], fx["x"].flip(0)], dim=2).contiguous()
        m.eval()
        with torch.no_grad():
            y_inf = m({"data": xe})["data"]
        m.train()
        y_tr = m({"data": xe})["data"]
        assert y_tr.requires_grad and y_tr.shape == y_inf.shape
        msgs.append((f"model_ensemble2_{kind}", ((y_tr - y_inf).abs().max() / y_inf.abs().max()).item(), 0.0))
    # two datasets (two encoders summed into one latent, two decoders): training forward == inference forward (pinned to the reference golden),
    # and every parameter receives a gradient
    from test_model_glue import build_two_dataset_model

    fx2 = torch.load(os.path.join(gdir, "model_forward_two_datasets.pt"), weights_only=False)
    m = build_two_dataset_model(fx2)
    m.load_state_dict(fx2["sd"], strict=True)
    m.train()
    y = m({k: v.clone() for k, v in fx2["x"].items()})
    sum(v.square().sum() for v in y.values()).backward()
    missing = [k for k, p in m.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:4]
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    msgs.append(("model_two_datasets", max(((y[k] - ref).abs().max() / ref.abs().max()).item() for k, ref in fx2["y"].items()), 0.0))
    return msgs


import pytest  # noqa: E402


@pytest.mark.parametrize("fn_name", ["_check", "_check_model", "_check_checkpoint", "_check_degenerate", "_check_random_configs"])
def test_sharded_training_step_matches_single_rank_gloo(fn_name):
    world = 2
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdv"), ret, fn_name), nprocs=world, join=True)
        for r in range(world):
            assert isinstance(ret.get(r), list), ret.get(r)
            for kind, err_x, err_p in ret[r]:
                assert err_x <= 2e-5 and err_p <= 5e-5, f"rank {r} {kind}: dx {err_x:.3e} dparams {err_p:.3e}"


# ---- the differentiable composition (layers/_train.py) against the gradient fixtures of the UNMODIFIED reference, on CPU ----------------
def _rel(a, b, floor=1e-6):
    a, b = a.detach().float(), b.detach().float()
    assert a.shape == b.shape, (tuple(a.shape), tuple(b.shape))
    return (a - b).abs().max().item() / max(b.abs().max().item(), floor)


def _param_errs(m, golden_grads):
    got = {n: p.grad for n, p in m.named_parameters()}
    floor = 1e-3 * max(g.abs().max().item() for g in golden_grads["params"].values())
    worst = 0.0
    for n, g in golden_grads["params"].items():
        assert got.get(n) is not None, f"no gradient for {n}"
        worst = max(worst, _rel(got[n], g, floor))
    return worst


def _check_fixtures(rank, world):
    """Forward and every gradient of the drop-in modules in training mode (stand-in arithmetic, fp32, CPU) against tests/golden/grads.pt and
    grads_r2.pt (PyTorch autograd of the unmodified reference modules): processors (plain, qk_norm, the four gatings, ConditionalLayerNorm),
    all four mappers (+ the ConditionalLayerNorm forward mapper)."""
    import anemoi_core_b200.layers as L
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g1 = torch.load(os.path.join(gdir, "grads.pt"), weights_only=False)
    g2 = torch.load(os.path.join(gdir, "grads_r2.pt"), weights_only=False)
    out = []

    def lk(c):
        return {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": c["condition_shape"],
                              "zero_init": False}}  # fmt: skip

    procs = [(g1, n) for n in ("gt_processor", "gt_processor_qknorm", "gnn_processor")]
    procs += [(g2, n) for n in ("gt_processor_glu", "gt_processor_swiglu", "gt_processor_geglu", "gt_processor_reglu", "gnn_processor_swiglu",
                                "gnn_processor_geglu", "gt_processor_condln")]  # fmt: skip
    for src, name in procs:
        c = src[name]
        cls = L.GNNProcessor if name.startswith("gnn") else L.GraphTransformerProcessor
        kw = {"layer_kernels": lk(c)} if name.endswith("condln") else {}
        m = cls(**kw, **c["cfg"])
        m.load_state_dict(c["sd"], strict=True)
        m.train()
        x, ea = c["x"].clone().requires_grad_(), c["edge_attr"].clone().requires_grad_()
        extra = {}
        if name.endswith("condln"):
            extra["cond"] = c["cond"].clone().requires_grad_()
        y = m(x, 1, GraphShardInfo(nodes=[x.shape[0]]) if name.startswith("gnn") else GraphShardInfo(), ea, c["edge_index"], **extra)
        (y * c["w"]).sum().backward()
        errs = [_rel(y, c["y"]), _rel(x.grad, c["grads"]["x"]), _rel(ea.grad, c["grads"]["edge_attr"]), _param_errs(m, c["grads"])]
        if extra:
            errs.append(_rel(extra["cond"].grad, c["grads"]["cond"]))
        out.append((name, max(errs)))
    mappers = [(g1, "gt_forward_mapper", L.GraphTransformerForwardMapper), (g1, "gt_backward_mapper", L.GraphTransformerBackwardMapper),
               (g1, "gnn_forward_mapper", L.GNNForwardMapper), (g1, "gnn_backward_mapper", L.GNNBackwardMapper),
               (g2, "gt_forward_mapper_condln", L.GraphTransformerForwardMapper)]  # fmt: skip
    for src, name, cls in mappers:
        c = src[name]
        cond = name.endswith("condln")
        m = cls(**({"layer_kernels": lk(c)} if cond else {}), **c["cfg"])
        m.load_state_dict(c["sd"], strict=True)
        m.train()
        xs, xd, ea = (c[k].clone().requires_grad_() for k in ("x_src", "x_dst", "edge_attr"))
        extra, conds = {}, ()
        if cond:
            conds = (c["cond_src"].clone().requires_grad_(), c["cond_dst"].clone().requires_grad_())
            extra["cond"] = conds
        res = m((xs, xd), 1, BipartiteGraphShardInfo(), ea, c["edge_index"], **extra)
        if name == "gnn_forward_mapper" or cond:
            ys = [res[1], res[0]]
        elif name == "gt_forward_mapper":
            ys = [res[1]]
        else:
            ys = [res]
        errs = [_rel(y, yr) for y, yr in zip(ys, c["y"])]
        sum((y * w).sum() for y, w in zip(ys, c["w"])).backward()
        for k, t in (("x_src", xs), ("x_dst", xd), ("edge_attr", ea)) + ((("cond_src", conds[0]), ("cond_dst", conds[1])) if cond else ()):
            if c["grads"][k] is None:
                assert t.grad is None or t.grad.abs().max().item() == 0.0
            else:
                errs.append(_rel(t.grad, c["grads"][k]))
        errs.append(_param_errs(m, c["grads"]))
        out.append((name, max(errs)))
    return out


def test_model_training_matches_reference_gradient_fixture_cpu():
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(1, os.path.join(d, "rdv"), ret, "_check_model_fixture"), nprocs=1, join=True)
        assert isinstance(ret.get(0), list), ret.get(0)
        for name, err_x, err_p in ret[0]:
            assert err_x <= 1e-4 and err_p <= 1e-4, f"{name}: output / input gradient {err_x:.3e}, parameter gradients {err_p:.3e}"
        print([(n, f"{a:.1e}", f"{b:.1e}") for n, a, b in ret[0]])


def test_training_composition_matches_reference_gradient_fixtures_cpu():
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(1, os.path.join(d, "rdv"), ret, "_check_fixtures"), nprocs=1, join=True)
        assert isinstance(ret.get(0), list), ret.get(0)
        for name, err in ret[0]:
            assert err <= 1e-4, f"{name}: max relative error {err:.3e}"
        print([(n, f"{e:.1e}") for n, e in ret[0]])
