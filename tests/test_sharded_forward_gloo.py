"""Sharded forward == single-rank forward under Gloo at world sizes 2, 3 and 4 (uneven shards), on CPU, with ``tests/_cpu_ops.py`` standing in
for the CUDA entry points: this exercises the HOST logic of the N > 1 path end to end - balanced node shards, 1-hop edge sharding (by the
processor or pre-sharded by the graph provider), the halo plan / exchange of the GraphTransformer processor (incl. qk_norm), the all-gather
path of the GNN processor, the sharded mappers of the whole ``EncProcDec`` step and the model-level ``AnemoiModelEncProcDec.forward``.
The kernels are covered by the ``-m gpu`` tests; the 2-GPU NCCL run of the same code is ``tests/test_gpu_multi.py``."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, init_file, fn_name, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        import _cpu_ops

        _cpu_ops.install()
        torch.set_grad_enabled(False)
        torch.set_num_threads(2)
        globals()[fn_name](rank, world)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


def run_distributed(fn_name, world):
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdv"), fn_name, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def _graph(n_src, n_dst, e, d, seed):
    g = torch.Generator().manual_seed(seed)
    dst = torch.cat([torch.arange(n_dst), torch.randint(0, n_dst, (e - n_dst,), generator=g)])
    ei = torch.stack([torch.randint(0, n_src, (e,), generator=g), dst])
    ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
    return ei, torch.randn(e, d, generator=g)


def _close(a, b, what):
    err = ((a - b).abs().max() / b.abs().max()).item()
    assert a.shape == b.shape and err <= 2e-5, f"{what}: {err:.3e}"


def check_processors(rank, world):
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    n, e, d = 101, 620, 5  # 101 nodes: uneven shards at every world size
    ei, ea = _graph(n, n, e, d, seed=1)
    sizes = get_balanced_partition_sizes(n, world)
    group = dist.group.WORLD
    for kind in ("gt", "gt_qknorm", "gnn"):
        torch.manual_seed(0)
        if kind.startswith("gt"):
            m = GraphTransformerProcessor(num_layers=2, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d, qk_norm=kind == "gt_qknorm")
            c = 64
        else:
            m = GNNProcessor(num_channels=32, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d)
            c = 32
        m.eval()
        x = torch.randn(n, c, generator=torch.Generator().manual_seed(2))
        full = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
        local = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
        assert local.shape[0] == sizes[rank]
        _close(gather_rows(local, sizes, group), full, kind)
        if kind == "gnn":  # opt-in halo form of the sharded GNN processor (node projections over local + halo rows only)
            import anemoi_core_b200.layers.processor as proc_mod

            proc_mod.GNN_HALO = True
            local_h = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
            proc_mod.GNN_HALO = False
            _close(gather_rows(local_h, sizes, group), full, "gnn, halo exchange")
        # bf16 inputs select the tensor-core host path (LayerNorm folded into the GEMMs, row statistics handed over as tensor tags,
        # K-padded operands): its bookkeeping must survive the sharding too.  Loose tolerance: every stage rounds to bf16.
        xb = x.bfloat16()
        ea16 = ea.bfloat16() if kind == "gnn" else ea  # (the GNN promotes x with its edge features; the GraphTransformer keeps attributes fp32)
        full16 = m(xb, 1, GraphShardInfo(nodes=[n]), ea16, ei)
        local16 = m(shard_rows(xb, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea16, ei, group)
        got16 = gather_rows(local16, sizes, group)
        assert full16.dtype == torch.bfloat16 and got16.dtype == torch.bfloat16
        rel = ((got16.float() - full16.float()).norm() / full16.float().norm()).item()
        assert rel <= 2e-2, f"{kind} bf16 host path: {rel:.3e}"
        assert ((full16.float() - full).norm() / full.norm()).item() <= 3e-2


def check_degenerate_graphs(rank, world):
    """Shapes a real mesh never has but a drop-in must survive: destinations without edges, ranks whose rows receive no edge at all, every source
    on one rank, a single edge, one node per rank.  Sharded == single rank for both processors (GraphTransformer: halo exchange; GNN: all-gather
    and halo form)."""
    import anemoi_core_b200.layers.processor as proc_mod
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    group = dist.group.WORLD
    d = 4
    g = torch.Generator().manual_seed(11)

    def graph(n, src, dst):
        ei = torch.stack([src, dst])
        ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
        return ei, torch.randn(ei.shape[1], d, generator=torch.Generator().manual_seed(int(ei.sum()) + n))

    cases = {}
    n = 2 * world + 1
    sz = get_balanced_partition_sizes(n, world)
    cases["sparse: destinations without edges"] = (n, *graph(n, torch.randint(0, n, (n,), generator=g), torch.randint(0, n, (n,), generator=g)))
    cases["all edges into the first rank's rows"] = (n, *graph(n, torch.randint(0, n, (3 * n,), generator=g), torch.randint(0, sz[0], (3 * n,), generator=g)))
    cases["all sources on the last rank"] = (n, *graph(n, torch.randint(n - sz[-1], n, (3 * n,), generator=g), torch.randint(0, n, (3 * n,), generator=g)))
    cases["a single edge"] = (n, *graph(n, torch.tensor([n - 1]), torch.tensor([0])))
    cases["one node per rank"] = (world, *graph(world, torch.randint(0, world, (2 * world,), generator=g), torch.randint(0, world, (2 * world,), generator=g)))
    for what, (n, ei, ea) in cases.items():
        sizes = get_balanced_partition_sizes(n, world)
        for kind in ("gt", "gnn"):
            torch.manual_seed(5)
            if kind == "gt":
                m, c = GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d).eval(), 32
            else:
                m, c = GNNProcessor(num_channels=16, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d).eval(), 16
            x = torch.randn(n, c, generator=torch.Generator().manual_seed(6))
            full = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
            for halo in ((False,) if kind == "gt" else (False, True)):
                proc_mod.GNN_HALO = halo
                local = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
                proc_mod.GNN_HALO = False
                assert local.shape[0] == sizes[rank]
                _close(gather_rows(local, sizes, group), full, f"{what}, {kind}{', halo form' if halo else ''}")


def check_derived_weights_follow_the_parameters(rank, world):
    """The derived weights of the inference path (casts, concatenations, LayerNorm / lin_edge folds; ``WeightPack``) against every way parameters
    change: an optimizer-style in-place update (version counter), ``load_state_dict``, an edit through ``.data`` (NO version bump: caught by the
    train() / eval() switch or by ``invalidate_packed_weights``); a frozen pack is left alone."""
    import copy

    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor
    from anemoi_core_b200.layers import _functional as Fn

    n, e, d = 40, 160, 5
    ei, ea = _graph(n, n, e, d, seed=21)
    for kind in ("gt", "gnn"):
        torch.manual_seed(3)
        if kind == "gt":
            m, c = GraphTransformerProcessor(num_layers=2, num_channels=32, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d).eval(), 32
        else:
            m, c = GNNProcessor(num_channels=48, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d).eval(), 48  # 48: the decomposed GraphConv form
        x = torch.randn(n, c, generator=torch.Generator().manual_seed(22))
        run = lambda mod: mod(x, 1, GraphShardInfo(nodes=[n]), ea, ei)  # noqa: E731

        def fresh():  # a new module with the current parameters: nothing derived is cached in it
            f = copy.deepcopy(m)
            Fn.invalidate_packed_weights(f)
            return run(f.eval())

        y0 = run(m)
        gen = torch.Generator().manual_seed(23)
        for p in m.parameters():  # optimizer-style update
            p.add_(0.05 * torch.randn(p.shape, generator=gen))
        y1 = run(m)
        assert (y1 - y0).abs().max() > 1e-3
        _close(y1, fresh(), f"{kind}: after an in-place parameter update")
        sd = {k: v + 0.05 * torch.randn(v.shape, generator=gen) for k, v in m.state_dict().items()}
        m.load_state_dict(sd, strict=True)
        _close(run(m), fresh(), f"{kind}: after load_state_dict")
        before = run(m)
        for p in m.parameters():  # through .data: neither the storage pointer nor the version counter moves
            p.data.add_(0.05 * torch.randn(p.shape, generator=gen))
        m.train()
        m.eval()
        y2 = run(m)
        assert (y2 - before).abs().max() > 1e-3
        _close(y2, fresh(), f"{kind}: .data edit, then train() / eval()")
        for p in m.parameters():
            p.data.add_(0.05 * torch.randn(p.shape, generator=gen))
        Fn.invalidate_packed_weights(m)
        _close(run(m), fresh(), f"{kind}: .data edit, then invalidate_packed_weights")
        Fn.freeze_packed_weights(m)
        held = [len(mod._pack._store) for mod in m.modules() if isinstance(getattr(mod, "_pack", None), Fn.WeightPack)]
        m.train()
        m.eval()
        assert held == [len(mod._pack._store) for mod in m.modules() if isinstance(getattr(mod, "_pack", None), Fn.WeightPack)] and sum(held) > 0
        Fn.freeze_packed_weights(m, False)


def check_inference_mode(rank, world):
    """``torch.inference_mode()`` (Lightning's validate / test / predict loops enter it by default): tensors created inside it track no version
    counter and ``t._version`` raises, so every identity-keyed cache (CSR plans, edge splits, halo plans, derived weights, row-statistics tags, graph
    provider outputs) must cope.  Processors (also sharded), all four mappers and the whole model: same result as under ``no_grad``."""
    import importlib.util

    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNBackwardMapper
    from anemoi_core_b200.layers import GNNForwardMapper
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerBackwardMapper
    from anemoi_core_b200.layers import GraphTransformerForwardMapper
    from anemoi_core_b200.layers import GraphTransformerProcessor

    group = dist.group.WORLD
    n, e, d = 41, 170, 5
    ei, ea = _graph(n, n, e, d, seed=31)
    sizes = get_balanced_partition_sizes(n, world)
    for kind, c in (("gt", 32), ("gnn", 16), ("gnn", 48)):
        torch.manual_seed(3)
        m = (GraphTransformerProcessor(num_layers=2, num_channels=c, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d, qk_norm=True) if kind == "gt"
             else GNNProcessor(num_channels=c, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d)).eval()  # fmt: skip
        x = torch.randn(n, c, generator=torch.Generator().manual_seed(32))
        ref = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
        with torch.inference_mode():
            xi, eai, eii = x.clone(), ea.clone(), ei.clone()  # inference tensors
            assert xi.is_inference() and eii.is_inference()
            _close(m(xi, 1, GraphShardInfo(nodes=[n]), eai, eii), ref, f"{kind} {c}: inference_mode")
            _close(m(xi, 1, GraphShardInfo(nodes=[n]), eai, eii), ref, f"{kind} {c}: inference_mode, cached plans")
            if world > 1:
                local = m(shard_rows(xi, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), eai, eii, group)
                _close(gather_rows(local, sizes, group), ref, f"{kind} {c}: inference_mode, sharded")
        if world == 1:
            # a validation phase under inference_mode followed by a training step on the SAME (normal) graph tensors: whatever the first phase
            # cached for this graph (plans built from tensors created inside inference mode) must serve the differentiable path as well
            import copy

            m2 = copy.deepcopy(m)
            with torch.inference_mode():
                _close(m(x, 1, GraphShardInfo(nodes=[n]), ea, ei), ref, f"{kind} {c}: inference_mode on normal tensors")
            grads = []
            for mod, graph in ((m, (ea, ei)), (m2, (ea.clone(), ei.clone()))):  # m2 on fresh graph tensors: nothing cached
                with torch.enable_grad():
                    mod.train()
                    xt = x.clone().requires_grad_()
                    mod(xt, 1, GraphShardInfo(nodes=[n]), *graph).square().sum().backward()
                    grads.append([xt.grad] + [p.grad for p in mod.parameters() if p.grad is not None])
                    mod.eval()
            assert len(grads[0]) == len(grads[1]) and all(torch.equal(a, b) for a, b in zip(*grads)), f"{kind} {c}: training after an inference_mode phase"
    if world > 1:
        return
    n_src, n_dst = 53, 37
    bi = BipartiteGraphShardInfo()
    g = torch.Generator().manual_seed(33)
    for cls, kw, cin in ((GraphTransformerForwardMapper, dict(in_channels_src=7, in_channels_dst=5, hidden_dim=32, num_heads=4, mlp_hidden_ratio=2), (7, 5)),
                         (GraphTransformerBackwardMapper, dict(in_channels_src=32, in_channels_dst=7, hidden_dim=32, out_channels_dst=4, num_heads=4, mlp_hidden_ratio=2), (32, 7)),
                         (GNNForwardMapper, dict(in_channels_src=7, in_channels_dst=5, hidden_dim=16, mlp_extra_layers=0), (7, 5)),
                         (GNNBackwardMapper, dict(in_channels_src=16, in_channels_dst=16, hidden_dim=16, out_channels_dst=4, mlp_extra_layers=0), (16, 16))):  # fmt: skip
        torch.manual_seed(4)
        m = cls(num_chunks=1, edge_dim=d, **kw).eval()
        mei, mea = _graph(n_src, n_dst, 150, d, seed=34)
        xs, xd = torch.randn(n_src, cin[0], generator=g), torch.randn(n_dst, cin[1], generator=g)
        ref = m((xs, xd), 1, bi, mea, mei)
        with torch.inference_mode():
            got = m((xs.clone(), xd.clone()), 1, bi, mea.clone(), mei.clone())
        for a, b in zip(got if isinstance(got, tuple) else (got,), ref if isinstance(ref, tuple) else (ref,)):
            _close(a, b, f"{cls.__name__}: inference_mode")
    spec = importlib.util.spec_from_file_location("_glue", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_model_glue.py"))
    glue = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(glue)
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_forward.pt"), weights_only=False)
    for kind in ("graphtransformer", "gnn"):
        m = glue.build_model(fx, kind)
        m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
        with torch.inference_mode():
            y = m({"data": fx["x"].clone()})["data"]
            y_again = m({"data": fx["x"].clone()})["data"]
        _close(y, fx["cases"][kind]["y"], f"AnemoiModelEncProcDec {kind}: inference_mode")
        assert torch.equal(y, y_again)


def check_random_configs(rank, world):
    """Random constructor configurations through the SHARDED forward (both strategies of the GraphTransformer processor, the GNN processor in its
    all-gather and halo forms; gated MLPs, qk_norm, edge_pre_mlp, attn_channels, ConditionalLayerNorm with a sharded conditioning tensor, several
    layers / chunks): sharded == single rank.  The same generator runs on every rank."""
    import anemoi_core_b200.layers.processor as proc_mod
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor

    group = dist.group.WORLD
    g = torch.Generator().manual_seed(101)

    def pick(options):
        return options[int(torch.randint(0, len(options), (1,), generator=g))]

    for case in range(14):
        n = int(torch.randint(3 * world, 90, (1,), generator=g))
        d = pick([3, 5, 11])
        ei, ea = _graph(n, n, int(torch.randint(n, 5 * n, (1,), generator=g)), d, seed=200 + case)
        sizes = get_balanced_partition_sizes(n, world)
        kw_call, cond = {}, None
        torch.manual_seed(300 + case)
        if case % 2 == 0:
            heads = pick([h for h in (2, 4, 8) if h % world == 0] or [world])
            c = heads * pick([8, 16])
            layers = pick([1, 2, 3])
            cfg = dict(num_layers=layers, num_channels=c, num_chunks=pick([k for k in (1, 2, 3) if layers % k == 0]), num_heads=heads, mlp_hidden_ratio=pick([2, 4]),
                       edge_dim=d, qk_norm=pick([False, True]), mlp_implementation=pick(["mlp", "glu", "swiglu", "geglu", "reglu"]),
                       shard_strategy=pick(["edges", "edges", "heads"]))  # fmt: skip
            if pick([False, True]):
                cfg["edge_pre_mlp"] = True
            if pick([False, True]):
                cfg["attn_channels"] = 2 * c
            if pick([False, False, True]):
                dc = 8
                cfg["layer_kernels"] = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": dc,
                                                      "zero_init": False}}  # fmt: skip
                cond = torch.randn(n, dc, generator=g)
            m = GraphTransformerProcessor(**cfg).eval()
            what = f"case {case}: GraphTransformerProcessor {cfg}"
        else:
            c, layers = pick([16, 32, 48]), pick([1, 2, 3])
            cfg = dict(num_channels=c, num_layers=layers, num_chunks=pick([k for k in (1, 2, 3) if layers % k == 0]), mlp_extra_layers=pick([0, 1]), edge_dim=d,
                       mlp_implementation=pick(["mlp", "swiglu"]))  # fmt: skip
            m = GNNProcessor(**cfg).eval()
            proc_mod.GNN_HALO = pick([False, True])
            what = f"case {case}: GNNProcessor {cfg} halo={proc_mod.GNN_HALO}"
        with torch.no_grad():
            for p in m.parameters():  # LayerNorm affine terms away from (1, 0), conditional scales away from 0
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
        x = torch.randn(n, c, generator=g)
        try:
            if cond is not None:
                full = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei, cond=cond)
                local = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group, cond=shard_rows(cond, sizes, group).contiguous())
            else:
                full = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
                local = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
        finally:
            proc_mod.GNN_HALO = False
        assert local.shape[0] == sizes[rank], what
        _close(gather_rows(local, sizes, group), full, what)


def check_heads_strategy(rank, world):
    """shard_strategy="heads" (Ulysses, block.py:689-759): nodes sharded outside the attention, heads inside; full edge list on every rank."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GraphTransformerProcessor

    n, e, d = 101, 620, 5
    ei, ea = _graph(n, n, e, d, seed=7)
    sizes = get_balanced_partition_sizes(n, world)
    group = dist.group.WORLD
    for qk_norm in (False, True):
        torch.manual_seed(3)
        kw = dict(num_layers=2, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d, qk_norm=qk_norm)
        single = GraphTransformerProcessor(**kw).eval()
        heads = GraphTransformerProcessor(shard_strategy="heads", **kw).eval()
        heads.load_state_dict(single.state_dict(), strict=True)
        x = torch.randn(n, 64, generator=torch.Generator().manual_seed(4))
        full = single(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
        local = heads(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
        _close(gather_rows(local, sizes, group), full, f"heads strategy qk_norm={qk_norm}")
        _close(heads(x, 1, GraphShardInfo(nodes=[n]), ea, ei), full, "heads-strategy module on one rank")


def check_heads_strategy_mappers(rank, world):
    """GraphTransformer mappers with shard_strategy="heads" (mapper.py:388-478): replicated sources (encoder) and sharded sources (decoder)."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.distributed.graph import shard_rows
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.layers import GraphTransformerBackwardMapper
    from anemoi_core_b200.layers import GraphTransformerForwardMapper

    n_grid, n_mesh, d = 83, 37, 4
    group = dist.group.WORLD
    gs, ms = get_balanced_partition_sizes(n_grid, world), get_balanced_partition_sizes(n_mesh, world)
    g = torch.Generator().manual_seed(8)
    xg, xm, lat = torch.randn(n_grid, 9, generator=g), torch.randn(n_mesh, 6, generator=g), torch.randn(n_mesh, 64, generator=g)
    for qk_norm in (False, True):
        kw = dict(hidden_dim=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d, qk_norm=qk_norm)
        ei, ea = _graph(n_grid, n_mesh, 260, d, seed=9)
        torch.manual_seed(5)
        enc = GraphTransformerForwardMapper(in_channels_src=9, in_channels_dst=6, **kw).eval()
        enc_h = GraphTransformerForwardMapper(in_channels_src=9, in_channels_dst=6, shard_strategy="heads", **kw).eval()
        enc_h.load_state_dict(enc.state_dict(), strict=True)
        _, full = enc((xg, xm), 1, BipartiteGraphShardInfo(), ea, ei)
        _, got = enc_h((xg, shard_rows(xm, ms, group).contiguous()), 1, BipartiteGraphShardInfo(src_nodes=None, dst_nodes=ms), ea, ei, group,
                       keep_x_dst_sharded=False)  # fmt: skip
        _close(got, full, f"forward mapper, heads strategy, qk_norm={qk_norm}")
        ei, ea = _graph(n_mesh, n_grid, 300, d, seed=10)
        torch.manual_seed(6)
        dec = GraphTransformerBackwardMapper(in_channels_src=64, in_channels_dst=9, out_channels_dst=5, **kw).eval()
        dec_h = GraphTransformerBackwardMapper(in_channels_src=64, in_channels_dst=9, out_channels_dst=5, shard_strategy="heads", **kw).eval()
        dec_h.load_state_dict(dec.state_dict(), strict=True)
        full = dec((lat, xg), 1, BipartiteGraphShardInfo(), ea, ei)
        got = dec_h((shard_rows(lat, ms, group).contiguous(), shard_rows(xg, gs, group).contiguous()), 1,
                    BipartiteGraphShardInfo(src_nodes=ms, dst_nodes=gs), ea, ei, group, keep_x_dst_sharded=False)  # fmt: skip
        _close(got, full, f"backward mapper, heads strategy, qk_norm={qk_norm}")


def check_enc_proc_dec(rank, world):
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.model import EncProcDec

    n_grid, n_mesh, d = 83, 37, 4
    gr = {}
    gr["enc_index"], gr["enc_attr"] = _graph(n_grid, n_mesh, 260, d, seed=3)
    gr["proc_index"], gr["proc_attr"] = _graph(n_mesh, n_mesh, 230, d, seed=4)
    gr["dec_index"], gr["dec_attr"] = _graph(n_mesh, n_grid, 300, d, seed=5)
    torch.manual_seed(1)
    model = EncProcDec("graphtransformer", in_grid=9, in_mesh=6, out_grid=5, num_channels=64, num_layers=2, edge_dim=d, num_heads=4).eval()
    g = torch.Generator().manual_seed(6)
    xg, xm = torch.randn(n_grid, 9, generator=g), torch.randn(n_mesh, 6, generator=g)
    full = model(xg, xm, gr)
    got = model(xg, xm, gr, dist.group.WORLD, get_balanced_partition_sizes(n_mesh, world), get_balanced_partition_sizes(n_grid, world))
    _close(got, full, "EncProcDec graphtransformer")
    torch.manual_seed(2)
    gnn = EncProcDec("gnn", in_grid=9, in_mesh=6, out_grid=5, num_channels=32, num_layers=2, edge_dim=d).eval()
    full = gnn(xg, xm, gr)
    got = gnn(xg, xm, gr, dist.group.WORLD, get_balanced_partition_sizes(n_mesh, world), None)  # GNN mappers replicated, processor sharded
    _close(got, full, "EncProcDec gnn")
    got = gnn(xg, xm, gr, dist.group.WORLD, get_balanced_partition_sizes(n_mesh, world), get_balanced_partition_sizes(n_grid, world))
    _close(got, full, "EncProcDec gnn, sharded GNN mappers")
    # the mappers alone, replicated inputs cut inside (ensure_sharded) and edges pre-sharded by nobody
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo

    src_full, dst_full = gnn.encoder((xg, xm), 1, BipartiteGraphShardInfo(), gr["enc_attr"], gr["enc_index"])
    src_l, dst_g = gnn.encoder((xg, xm), 1, BipartiteGraphShardInfo(), gr["enc_attr"], gr["enc_index"], dist.group.WORLD, keep_x_dst_sharded=False)
    _close(dst_g, dst_full, "GNNForwardMapper dst (gathered)")
    _close(gather_rows(src_l, get_balanced_partition_sizes(n_grid, world), dist.group.WORLD), src_full, "GNNForwardMapper src shard")


def check_degenerate_enc_proc_dec(rank, world):
    """The whole sharded step on bipartite graphs a real grid / mesh pair never has: mesh rows no grid point maps to, a rank whose mesh rows get no
    encoder edge, every encoder source on the last rank, every decoder edge into the first rank's grid rows, grid points the decoder never writes."""
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.model import EncProcDec

    n_grid, n_mesh, d = 11 * world + 3, 4 * world + 1, 4
    gs, ms = get_balanced_partition_sizes(n_grid, world), get_balanced_partition_sizes(n_mesh, world)
    g = torch.Generator().manual_seed(13)

    def graph(src, dst):
        ei = torch.stack([src, dst])
        ei = ei[:, torch.sort(ei[1], stable=True)[1]].contiguous()
        return ei, torch.randn(ei.shape[1], d, generator=g)

    def rnd(lo, hi, k):
        return torch.randint(lo, hi, (k,), generator=g)

    variants = {
        "sparse": dict(enc=graph(rnd(0, n_grid, n_mesh), rnd(0, n_mesh, n_mesh)), proc=graph(rnd(0, n_mesh, n_mesh), rnd(0, n_mesh, n_mesh)),
                       dec=graph(rnd(0, n_mesh, n_grid // 2), rnd(0, n_grid, n_grid // 2))),
        "one-sided": dict(enc=graph(rnd(n_grid - gs[-1], n_grid, 40), rnd(0, ms[0], 40)), proc=graph(rnd(n_mesh - ms[-1], n_mesh, 30), rnd(0, ms[0], 30)),
                          dec=graph(rnd(n_mesh - ms[-1], n_mesh, 40), rnd(0, gs[0], 40))),
    }  # fmt: skip
    xg, xm = torch.randn(n_grid, 9, generator=g), torch.randn(n_mesh, 6, generator=g)
    g0, m0 = sum(gs[:rank]), sum(ms[:rank])
    for what, v in variants.items():
        gr = {"enc_index": v["enc"][0], "enc_attr": v["enc"][1], "proc_index": v["proc"][0], "proc_attr": v["proc"][1], "dec_index": v["dec"][0],
              "dec_attr": v["dec"][1]}  # fmt: skip
        torch.manual_seed(1)
        gt = EncProcDec("graphtransformer", in_grid=9, in_mesh=6, out_grid=5, num_channels=32, num_layers=2, edge_dim=d, num_heads=4).eval()
        full = gt(xg, xm, gr)
        _close(gt(xg, xm, gr, dist.group.WORLD, ms, gs), full, f"{what}: EncProcDec graphtransformer")
        part = gt(xg[g0 : g0 + gs[rank]].contiguous(), xm[m0 : m0 + ms[rank]].contiguous(), gr, model_comm_group=dist.group.WORLD, mesh_shards=ms,
                  grid_shards=gs, keep_output_sharded=True, inputs_sharded=True)  # fmt: skip
        _close(part, full[g0 : g0 + gs[rank]], f"{what}: EncProcDec graphtransformer, sharded inputs and outputs")
        torch.manual_seed(2)
        gnn = EncProcDec("gnn", in_grid=9, in_mesh=6, out_grid=5, num_channels=16, num_layers=2, edge_dim=d).eval()
        full = gnn(xg, xm, gr)
        _close(gnn(xg, xm, gr, dist.group.WORLD, ms, None), full, f"{what}: EncProcDec gnn")
        _close(gnn(xg, xm, gr, dist.group.WORLD, ms, gs), full, f"{what}: EncProcDec gnn, sharded mappers")


def check_model_forward(rank, world):
    """``AnemoiModelEncProcDec.forward`` with ``model_comm_group``: hidden rows sharded, provider-sharded edges (global dst ids), replicated grid."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_glue", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_model_glue.py"))
    glue = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(glue)
    fx = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_forward.pt"), weights_only=False)
    for kind in ("graphtransformer", "gnn"):
        m = glue.build_model(fx, kind)
        m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
        x = {"data": fx["x"][:1].contiguous()}  # batch 1: the only batch size a sharded model accepts (encoder_processor_decoder.py:165-183)
        full = m(x)["data"]
        got = m(x, model_comm_group=dist.group.WORLD)["data"]
        _close(got, full, f"AnemoiModelEncProcDec {kind} sharded")
        if world == 2:  # the single-rank result is the reference golden's first batch element
            ref = fx["cases"][kind]["y"][:1]
            assert ((full - ref).abs().max() / ref.abs().max()).item() <= 1e-4


def check_goldens_single_rank(rank, world):
    """The mirror classes' HOST logic (weight packing / folding, buffer layouts, edge preparation, call order) on one rank, stand-in arithmetic,
    against the goldens of the unmodified reference modules: catches host-side regressions in the CPU suite; the kernels are the GPU suite's job."""
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNBackwardMapper
    from anemoi_core_b200.layers import GNNForwardMapper
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerBackwardMapper
    from anemoi_core_b200.layers import GraphTransformerForwardMapper
    from anemoi_core_b200.layers import GraphTransformerProcessor

    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def load(name):
        return torch.load(os.path.join(gdir, name + ".pt"), weights_only=False)

    def strip(cfg):
        return {k: v for k, v in cfg.items()}

    for name in ("gt_processor_small", "gt_processor_qknorm", "gt_processor_edge_pre_mlp", "gt_processor_attn_channels", "gt_processor_swiglu",
                 "gt_processor_glu", "gt_processor_geglu", "gt_processor_reglu"):  # fmt: skip
        g = load(name)
        m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, **strip(g["cfg"])).eval()
        m.load_state_dict(g["sd"], strict=True)
        _close(m(g["x"], 1, GraphShardInfo(), g["edge_attr"], g["edge_index"]), g["y"], name)
    g = load("gt_processor_unsorted")
    m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, **strip(g["cfg"])).eval()
    m.load_state_dict(g["sd"], strict=True)
    _close(m(g["x"], 1, GraphShardInfo(), g["edge_attr"], g["edge_index"], None, edges_are_dst_sorted=False), g["y"], "gt_processor_unsorted")
    for name in ("gnn_processor_small", "gnn_processor_cfg1", "gnn_processor_extra_layers", "gnn_processor_swiglu"):
        g = load(name)
        m = GNNProcessor(num_chunks=1, **{"mlp_extra_layers": 0, **strip(g["cfg"])}).eval()
        m.load_state_dict(g["sd"], strict=True)
        _close(m(g["x"], 1, GraphShardInfo(nodes=[g["x"].shape[0]]), g["edge_attr"], g["edge_index"]), g["y"], name)
    bi = BipartiteGraphShardInfo()
    g = load("gnn_forward_mapper")
    m = GNNForwardMapper(num_chunks=1, mlp_extra_layers=0, **g["cfg"]).eval()
    m.load_state_dict(g["sd"], strict=True)
    ys, yd = m((g["x_src"], g["x_dst"]), 1, bi, g["edge_attr"], g["edge_index"])
    _close(ys, g["y_src"], "gnn_forward_mapper src"), _close(yd, g["y_dst"], "gnn_forward_mapper dst")
    g = load("gnn_backward_mapper")
    m = GNNBackwardMapper(num_chunks=1, mlp_extra_layers=0, **g["cfg"]).eval()
    m.load_state_dict(g["sd"], strict=True)
    _close(m((g["x_src"], g["x_dst"]), 1, bi, g["edge_attr"], g["edge_index"]), g["y"], "gnn_backward_mapper")
    for name in ("gt_forward_mapper_chunks1", "gt_forward_mapper_chunks4"):
        g = load(name)
        m = GraphTransformerForwardMapper(mlp_hidden_ratio=4, **g["cfg"]).eval()
        m.load_state_dict(g["sd"], strict=True)
        _close(m((g["x_src"], g["x_dst"]), 1, bi, g["edge_attr"], g["edge_index"])[1], g["y_dst"], name)
    g = load("gt_backward_mapper")
    m = GraphTransformerBackwardMapper(mlp_hidden_ratio=4, **g["cfg"]).eval()
    m.load_state_dict(g["sd"], strict=True)
    _close(m((g["x_src"], g["x_dst"]), 1, bi, g["edge_attr"], g["edge_index"]), g["y"], "gt_backward_mapper")
    g = load("gt_processor_condln")  # ConditionalLayerNorm kernels + cond= (processor and forward mapper)
    lk = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": g["condition_shape"],
                        "zero_init": False}}  # fmt: skip
    m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **g["cfg"]).eval()
    m.load_state_dict(g["sd"], strict=True)
    _close(m(g["x"], 1, GraphShardInfo(), g["edge_attr"], g["edge_index"], cond=g["cond"]), g["y"], "gt_processor_condln")
    mp_ = g["mapper"]
    m = GraphTransformerForwardMapper(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **mp_["cfg"]).eval()
    m.load_state_dict(mp_["sd"], strict=True)
    _close(m((mp_["x_src"], mp_["x_dst"]), 1, bi, mp_["edge_attr"], mp_["edge_index"], cond=(mp_["cond_src"], mp_["cond_dst"]))[1], mp_["y_dst"],
           "gt_forward_mapper condln")  # fmt: skip
    # the whole model, one and two datasets (graph providers, node attributes, assembly, latent sum, residual, boundings)
    import importlib.util

    spec = importlib.util.spec_from_file_location("_glue", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_model_glue.py"))
    glue = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(glue)
    fx = load("model_forward")
    for kind in ("graphtransformer", "gnn"):
        m = glue.build_model(fx, kind)
        m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
        _close(m({"data": fx["x"]})["data"], fx["cases"][kind]["y"], f"model {kind}")
    fx = load("model_forward_two_datasets")
    m = glue.build_two_dataset_model(fx)
    m.load_state_dict(fx["sd"], strict=True)
    y = m(fx["x"])
    for name, ref in fx["y"].items():
        _close(y[name], ref, f"two-dataset model, {name}")


def test_host_logic_against_reference_goldens():
    run_distributed("check_goldens_single_rank", 1)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_processors(world):
    run_distributed("check_processors", world)


@pytest.mark.parametrize("world", [2, 3])
def test_degenerate_graphs(world):
    run_distributed("check_degenerate_graphs", world)


@pytest.mark.parametrize("world", [2, 3])
def test_degenerate_enc_proc_dec(world):
    run_distributed("check_degenerate_enc_proc_dec", world)


@pytest.mark.parametrize("world", [1, 2])
def test_inference_mode(world):
    run_distributed("check_inference_mode", world)


@pytest.mark.parametrize("world", [2, 4])
def test_random_configs_sharded(world):
    run_distributed("check_random_configs", world)


def test_derived_weights_follow_the_parameters():
    run_distributed("check_derived_weights_follow_the_parameters", 1)


@pytest.mark.parametrize("world", [2, 4])
def test_heads_strategy(world):
    run_distributed("check_heads_strategy", world)


@pytest.mark.parametrize("world", [2, 4])
def test_heads_strategy_mappers(world):
    run_distributed("check_heads_strategy_mappers", world)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_enc_proc_dec(world):
    run_distributed("check_enc_proc_dec", world)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_model_forward(world):
    run_distributed("check_model_forward", world)
