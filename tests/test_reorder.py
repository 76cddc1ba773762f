"""Locality plan (layers/_reorder.py): integer properties of the plan, the locality it buys on the icosahedral multi-scale mesh, and - with
the CPU stand-in arithmetic in a subprocess - that a processor run in the permuted order returns what the unpermuted run returns."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

from anemoi_core_b200.layers import _reorder as RO
from anemoi_core_b200.synthetic import icosphere_multiscale

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _mesh(level):
    xyz, ei = icosphere_multiscale(level)
    ei = torch.from_numpy(ei.astype(np.int64))
    return xyz, ei[:, torch.sort(ei[1], stable=True).indices].contiguous()  # dst-sorted, as the graph providers hand it over


def test_plan_is_a_consistent_relabelling():
    xyz, ei = _mesh(4)  # 2 562 nodes
    n = xyz.shape[0]
    plan = RO.locality_plan(ei, n)
    assert plan is not None and plan is RO.locality_plan(ei, n)  # cached on the tensor
    perm, rank = plan.perm.long(), plan.rank.long()
    assert torch.equal(torch.sort(perm).values, torch.arange(n)) and torch.equal(perm[rank], torch.arange(n))
    new = plan.edge_index
    assert bool((new[1][1:] >= new[1][:-1]).all())  # dst-sorted
    assert torch.equal(perm[new[0]], ei[0][plan.edge_perm]) and torch.equal(perm[new[1]], ei[1][plan.edge_perm])  # same edges, relabelled
    ea = torch.arange(ei.shape[1], dtype=torch.float32).view(-1, 1)
    assert torch.equal(RO.permute_edge_attr(ea, plan).view(-1).long(), plan.edge_perm)
    assert RO.locality_plan(ei[:, :0].contiguous(), n) is None and RO.locality_plan(_mesh(2)[1], 162) is None  # nothing to gain


def test_locality_on_the_icosahedral_mesh():
    """Topology alone recovers the sphere well enough: a tile of 16 destinations re-uses its sources > 2x (1.1x in the reference order)."""
    xyz, ei = _mesh(5)  # 10 242 nodes, 81 840 edges
    before = RO.source_reuse(ei, 16)
    plan = RO.locality_plan(ei, xyz.shape[0])
    after = RO.source_reuse(plan.edge_index, 16)
    with_coords = RO.source_reuse(RO.locality_plan(ei.clone(), xyz.shape[0], coords=xyz).edge_index, 16)
    assert before < 1.3 and after > 2.0 and with_coords > 2.0, (before, after, with_coords)
    d = (plan.edge_index[0] - plan.edge_index[1]).abs().float()
    d0 = (ei[0] - ei[1]).abs().float()
    assert d.median() * 8 < d0.median()


def _worker(rank, world, init_file, ret):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        import _cpu_ops

        _cpu_ops.install()
        torch.set_grad_enabled(False)
        from anemoi_core_b200.distributed.shapes import GraphShardInfo
        from anemoi_core_b200.layers import GNNProcessor
        from anemoi_core_b200.layers import GraphTransformerProcessor

        xyz, ei = _mesh(4)
        n, e, d = xyz.shape[0], ei.shape[1], 5
        g = torch.Generator().manual_seed(0)
        ea = torch.randn(e, d, generator=g)
        msgs = []
        for kind in ("gt", "gnn"):
            torch.manual_seed(1)
            if kind == "gt":
                m, c = GraphTransformerProcessor(num_layers=2, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=2, edge_dim=d).eval(), 64
            else:
                m, c = GNNProcessor(num_channels=32, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d).eval(), 32
            x = torch.randn(n, c, generator=g)
            RO.ENABLED = False
            ref = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
            RO.ENABLED = True
            got = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
            RO.ENABLED = False
            assert RO.locality_plan(ei, n) is not None
            msgs.append((kind, ((got - ref).abs().max() / ref.abs().max()).item()))
        ret[rank] = msgs
    except Exception as ex:  # noqa: BLE001
        import traceback

        ret[rank] = f"{type(ex).__name__}: {ex}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


def test_reordered_forward_equals_plain_forward():
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as dname:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(1, os.path.join(dname, "rdv"), ret), nprocs=1, join=True)
        assert isinstance(ret[0], list), ret[0]
        for kind, err in ret[0]:
            assert err <= 2e-5, (kind, err)
