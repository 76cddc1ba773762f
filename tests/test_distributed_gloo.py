"""world_size-2 Gloo tests (CPU) of the multi-GPU host logic: dst-range edge sharding (reference shard_edges_1hop,
khop_edges.py:266-314), balanced node shards, and the forward halves of the sharding collectives (gather / shard rows,
distributed/graph.py:66-135) with unequal shard sizes.  Kernels are not involved (no GPU here); the 2-GPU NCCL run of the
same code path is the `--gpus 2` bench / gpurun test."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes


def _worker(rank, world, init_file, fn_name, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        ret[rank] = "ok"
    except Exception as e:  # noqa: BLE001
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def run_distributed(fn_name, world=2):
    with tempfile.TemporaryDirectory() as d:
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdv"), fn_name, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def _graph(n=37, e=211, seed=3):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)])
    ei = ei[:, torch.sort(ei[1], stable=True)[1]]
    return ei, torch.arange(e, dtype=torch.float32).view(-1, 1)


def check_gather_and_shard(rank, world):
    from anemoi_core_b200.distributed.graph import gather_rows
    from anemoi_core_b200.distributed.graph import shard_rows

    full = torch.arange(37 * 3, dtype=torch.float32).view(37, 3)
    sizes = get_balanced_partition_sizes(37, world)  # [19, 18]: unequal -> list form of all_gather
    local = shard_rows(full, sizes, dist.group.WORLD)
    assert local.shape[0] == sizes[rank]
    again = gather_rows(local.clone(), sizes, dist.group.WORLD)
    assert torch.equal(again, full)
    even = torch.arange(40.0).view(20, 2)
    assert torch.equal(gather_rows(shard_rows(even, [10, 10], dist.group.WORLD).clone(), [10, 10], dist.group.WORLD), even)
    with pytest.raises(ValueError):
        gather_rows(local, None, dist.group.WORLD)


def check_edge_sharding(rank, world):
    from anemoi_core_b200.layers.processor import _shard_edges_by_dst

    ei, ea = _graph()
    n = 37
    sizes = get_balanced_partition_sizes(n, world)
    start = sum(sizes[:rank])
    for relabel in (False, True):
        ea_l, ei_l, edge_sizes = _shard_edges_by_dst(ea, ei, n, n, dist.group.WORLD, relabel_dst=relabel)
        # every edge whose dst lies in this rank's node range, in the original order, nothing else
        mask = (ei[1] >= start) & (ei[1] < start + sizes[rank])
        assert torch.equal(ea_l.view(-1), ea[mask].view(-1))
        assert torch.equal(ei_l[0], ei[0, mask])
        assert torch.equal(ei_l[1], ei[1, mask] - (start if relabel else 0))
        assert edge_sizes[rank] == int(mask.sum()) and sum(edge_sizes) == ei.shape[1]
        # all ranks agree on the split
        t = torch.tensor(edge_sizes)
        other = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(other, t)
        assert all(torch.equal(o, t) for o in other)
    # cached: same tensors back on the second call
    a1 = _shard_edges_by_dst(ea, ei, n, n, dist.group.WORLD, relabel_dst=True)[1]
    a2 = _shard_edges_by_dst(ea, ei, n, n, dist.group.WORLD, relabel_dst=True)[1]
    assert a1 is a2


def test_gather_and_shard_rows_world2():
    run_distributed("check_gather_and_shard", 2)


def test_edge_sharding_world2():
    run_distributed("check_edge_sharding", 2)


def test_edge_sharding_world3_uneven():
    run_distributed("check_edge_sharding", 3)


def check_halo_exchange(rank, world):
    """Halo plan + exchange (distributed/halo.py; reference block.py:1120-1183): after the exchange every local edge finds, at its relabelled
    source position of the compact table, exactly the row its GLOBAL source id names; only referenced remote rows travel."""
    from anemoi_core_b200.distributed.halo import halo_plan_for

    import _cpu_ops  # tests/: plain-PyTorch stand-in for the CUDA entry points (the exchange packs its rows with ops.cast_pad)

    _cpu_ops.install()
    from anemoi_core_b200.layers.processor import _shard_edges_by_dst

    for n, e, seed in ((37, 211, 3), (64, 40, 4), (50, 0, 5)):  # dense, sparse (small halos, some empty), no edges at all
        ei, ea = _graph(n, e, seed)
        sizes = get_balanced_partition_sizes(n, world)
        start = sum(sizes[:rank])
        _, ei_l, _ = _shard_edges_by_dst(ea, ei, n, n, dist.group.WORLD, relabel_dst=True)  # global src, local dst
        plan = halo_plan_for(ei_l, sizes, dist.group.WORLD)
        assert plan is halo_plan_for(ei_l, sizes, dist.group.WORLD)  # cached on the tensor
        full = torch.arange(n * 4, dtype=torch.float32).view(n, 4) * 0.5 + 1.0
        table = torch.full((plan.n_table, 4), float("nan"))
        table[: plan.n_local] = full[start : start + sizes[rank]]
        plan.exchange(table)
        assert plan.n_local == sizes[rank] and not torch.isnan(table).any()
        assert torch.equal(table[plan.edge_index[0]], full[ei_l[0]])  # every edge reads the right source row
        assert torch.equal(plan.edge_index[1], ei_l[1])
        # the halo is exactly the set of referenced remote sources, grouped by owner in ascending id order
        remote = torch.unique(ei_l[0][(ei_l[0] < start) | (ei_l[0] >= start + sizes[rank])])
        assert torch.equal(plan.halo_ids, remote) and plan.n_halo == remote.numel() and plan.recv_splits[rank] == 0
        assert torch.equal(table[plan.n_local :], full[plan.halo_ids])
        table2 = torch.full((plan.n_table, 4), float("nan"))  # split form used by the block (start / overlap / finish)
        table2[: plan.n_local] = table[: plan.n_local]
        plan.exchange_start(table2)
        plan.exchange_finish()
        assert torch.equal(table2, table)
        # what we send is what the others receive
        t = torch.tensor(plan.send_splits)
        others = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(others, torch.tensor(plan.recv_splits))
        assert [int(others[r][rank]) for r in range(world)] == plan.send_splits


def test_halo_exchange_world2():
    run_distributed("check_halo_exchange", 2)


def test_halo_exchange_world3_uneven():
    run_distributed("check_halo_exchange", 3)
