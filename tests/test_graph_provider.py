"""Graph provider / trainable tensors / named node attributes against golden vectors generated from the
unmodified reference classes (``oracle/gen_golden.py::graph_provider_cases``; layers/graph_provider.py:145-291, layers/graph.py:20-118):
bit-exact edge_index, exact attribute values, reference ``state_dict`` keys, identity-stable outputs (the caching contract), and the
1-hop edge sharding under a world-size-2 Gloo group."""
import os

import pytest
import torch

from anemoi_core_b200.distributed import khop_edges as K
from anemoi_core_b200.layers import NamedNodesAttributes
from anemoi_core_b200.layers import NoOpGraphProvider
from anemoi_core_b200.layers import StaticGraphProvider
from anemoi_core_b200.layers import TrainableTensor
from anemoi_core_b200.layers import create_graph_provider


def _provider(g):
    sub = {"edge_index": g["edge_index"], "edge_length": g["edge_length"], "edge_dirs": g["edge_dirs"]}
    p = StaticGraphProvider(graph=sub, edge_attributes=["edge_length", "edge_dirs"], src_size=g["n_src"], dst_size=g["n_dst"], trainable_size=3)
    missing, unexpected = p.load_state_dict(g["sd"], strict=True)
    assert not missing and not unexpected
    return p


@torch.no_grad()  # the forward / inference path; with grad enabled the trainable tensor is assembled differentiably and not cached
def test_static_graph_provider_matches_reference(golden):
    g = golden("graph_provider")
    p = _provider(g)
    assert sorted(p.state_dict().keys()) == sorted(g["sd"].keys())  # trainable.trainable + trainable_layout_version only
    assert p.edge_dim == g["edge_dim"] == 6
    for bs, ref in g["edges"].items():
        ea, ei, sizes = p.get_edges(batch_size=bs, model_comm_group=None)
        assert sizes is None
        assert ei.dtype == torch.int64 and torch.equal(ei, ref["edge_index"])  # bit-exact, dst-sorted, batch-expanded
        assert torch.equal(ea, ref["edge_attr"])
        assert K.is_edge_index_dst_sorted(ei[:, : g["edge_index"].shape[1]])
        # the caching contract: the same tensors come back, so downstream CSR plans keyed on identity are hits
        ea2, ei2, _ = p.get_edges(batch_size=bs, model_comm_group=None, shard_edges=False, act_checkpoint=False)
        assert ea2 is ea and ei2 is ei
    # a new version of the trainable parameter invalidates the attribute cache (and only that)
    ea, ei, _ = p.get_edges(batch_size=1)
    with torch.no_grad():
        p.trainable.trainable.add_(1.0)
    ea3, ei3, _ = p.get_edges(batch_size=1)
    assert ei3 is ei and ea3 is not ea
    assert torch.equal(ea3[:, :3], ea[:, :3]) and torch.allclose(ea3[:, 3:], g["edges"][1]["edge_attr"][:, 3:] + 1.0)


def test_trainable_tensor_grad_path_is_differentiable():
    t = TrainableTensor(tensor_size=5, trainable_size=2)
    x = torch.randn(5, 3)
    out = t(x, batch_size=2)
    assert out.shape == (10, 5) and out.requires_grad
    out.sum().backward()
    assert torch.equal(t.trainable.grad, torch.full((5, 2), 2.0))
    assert TrainableTensor(tensor_size=5, trainable_size=0).trainable is None
    assert torch.equal(TrainableTensor(5, 0)(x, 1), x)


def test_named_nodes_attributes_match_reference(golden):
    g = golden("graph_provider")
    a = NamedNodesAttributes({"hidden": 4}, {name: {"x": c} for name, c in g["coords"].items()})
    missing, unexpected = a.load_state_dict(g["attrs_sd"], strict=True)
    assert not missing and not unexpected
    assert a.attr_ndims == g["attr_ndims"] and a.num_nodes == g["num_nodes"]
    for (name, bs), ref in g["attrs"].items():
        with torch.no_grad():
            assert torch.equal(a(name, batch_size=bs), ref)
    for name, ref in g["coords_back"].items():
        assert torch.equal(a.get_coordinates(name), ref)


def test_create_graph_provider_factory(golden):
    g = golden("graph_provider")
    sub = {"edge_index": g["edge_index"], "edge_length": g["edge_length"], "edge_dirs": g["edge_dirs"]}
    assert isinstance(create_graph_provider(graph=sub, edge_attributes=["edge_length"], src_size=g["n_src"], dst_size=g["n_dst"]), StaticGraphProvider)
    noop = create_graph_provider(graph=None)
    assert isinstance(noop, NoOpGraphProvider) and noop.edge_dim == 0 and noop.get_edges(batch_size=2) == (None, None, None)


def _worker(rank, world, init_file, fixture, out):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)  # no fixed port: nothing to collide with
    try:
        g = torch.load(fixture, weights_only=False)
        p = _provider(g)
        torch.set_grad_enabled(False)
        ea, ei, sizes = p.get_edges(batch_size=1, model_comm_group=dist.group.WORLD)
        ea_b, ei_b, _ = p.get_edges(batch_size=1, model_comm_group=dist.group.WORLD)
        # what the processors / mappers do with the two edge forms must agree: pre-sharded edges (global dst ids) relabelled to the local
        # rows == the full list cut and relabelled by the processor itself
        from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
        from anemoi_core_b200.layers.processor import _localise_presharded_edges
        from anemoi_core_b200.layers.processor import _shard_edges_by_dst

        full_ea, full_ei, _ = p.get_edges(batch_size=1, model_comm_group=dist.group.WORLD, shard_edges=False)
        splits = get_balanced_partition_sizes(g["n_dst"], world)
        ea_own, ei_own, sizes_own = _shard_edges_by_dst(full_ea, full_ei, g["n_dst"], g["n_src"], dist.group.WORLD, relabel_dst=True)
        assert sizes_own == sizes and torch.equal(ea_own, ea)
        assert torch.equal(_localise_presharded_edges(ei, splits, dist.group.WORLD), ei_own)
        torch.save({"ea": ea.clone(), "ei": ei.clone(), "sizes": sizes, "stable": ea_b.data_ptr() == ea.data_ptr() and ei_b is ei}, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def test_sharded_edges_world2_gloo(golden, tmp_path):
    """shard_edges=True under a 2-rank group: contiguous 1-hop edge ranges of the dst-sorted list (khop_edges.py:266-314), global ids kept."""
    import torch.multiprocessing as mp

    g = golden("graph_provider")
    fixture = os.path.join(os.path.dirname(__file__), "golden", "graph_provider.pt")
    out = str(tmp_path / "shard")
    mp.spawn(_worker, args=(2, str(tmp_path / "rdv"), fixture, out), nprocs=2, join=True)
    full_ea, full_ei = g["edges"][1]["edge_attr"], g["edges"][1]["edge_index"]
    part = K.build_graph_partition(full_ei, 2, (g["n_src"], g["n_dst"]))
    got = [torch.load(f"{out}.{r}", weights_only=False) for r in range(2)]
    assert got[0]["sizes"] == got[1]["sizes"] == list(part.edge_splits)
    assert torch.equal(torch.cat([got[0]["ei"], got[1]["ei"]], 1), full_ei) and torch.equal(torch.cat([got[0]["ea"], got[1]["ea"]]), full_ea)
    for r in range(2):
        d0, d1 = part.dst_range(r)
        dst = got[r]["ei"][1]
        assert got[r]["stable"] and (dst.numel() == 0 or (int(dst.min()) >= d0 and int(dst.max()) < d1))
