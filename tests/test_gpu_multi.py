"""2-GPU (NCCL) parity of the dst-range sharded processors against the single-GPU result (-m gpu, needs >= 2 devices;
skipped on a 1-GPU box).  Reference behaviour: layers/block.py:375-391 (GNN: all-gather x, local edges, keep local rows)
and block.py:1148-1183 (GraphTransformer edges strategy).  Outputs must be identical to the unsharded run up to fp32
summation order (the per-row arithmetic is the same kernels on the same rows): tolerance 1e-5 relative."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, ret):
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
        from anemoi_core_b200.distributed.graph import gather_rows
        from anemoi_core_b200.distributed.graph import shard_rows
        from anemoi_core_b200.distributed.shapes import GraphShardInfo
        from anemoi_core_b200.layers import GNNProcessor
        from anemoi_core_b200.layers import GraphTransformerProcessor
        from anemoi_core_b200.synthetic import build_graph

        gr = build_graph("o32", mesh_level=4)
        n = gr["n_mesh"]
        ea, ei = gr["proc_attr"].cuda(), gr["proc_index"].cuda()
        sizes = get_balanced_partition_sizes(n, world)
        group = dist.group.WORLD
        msgs = []
        for kind in ("gt", "gt_qknorm", "gnn"):
            for dt in (torch.float32, torch.bfloat16):
                torch.manual_seed(0)
                if kind.startswith("gt"):
                    m = GraphTransformerProcessor(num_layers=2, num_channels=256, num_chunks=1, num_heads=8, mlp_hidden_ratio=4, edge_dim=gr["edge_dim"],
                                                  qk_norm=kind == "gt_qknorm")  # fmt: skip
                else:
                    m = GNNProcessor(num_channels=128, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=gr["edge_dim"])
                m = m.cuda().eval()
                c = 256 if kind.startswith("gt") else 128
                x = torch.randn(n, c, generator=torch.Generator().manual_seed(1)).cuda()
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
                    full = m(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
                    local = m(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
                    got = gather_rows(local, sizes, group)
                err = ((got.float() - full.float()).abs().max() / full.float().abs().max()).item()
                tol = 1e-5 if dt == torch.float32 else 2e-2
                msgs.append((kind, str(dt), err, err <= tol))
        # whole encoder -> processor -> decoder step, every stage dst-range sharded, against the single-GPU step
        from anemoi_core_b200.model import EncProcDec

        torch.manual_seed(5)
        model = EncProcDec("graphtransformer", in_grid=20, in_mesh=12, out_grid=9, num_channels=256, num_layers=2, edge_dim=gr["edge_dim"],
                           num_heads=8).cuda().eval()  # fmt: skip
        gd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in gr.items()}
        gen = torch.Generator().manual_seed(9)
        xg, xm = torch.randn(gr["n_grid"], 20, generator=gen).cuda(), torch.randn(n, 12, generator=gen).cuda()
        gsz = get_balanced_partition_sizes(gr["n_grid"], world)
        for dt in (torch.float32, torch.bfloat16):
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
                full = model(xg, xm, gd)
                got = model(xg, xm, gd, group, sizes, gsz)
            err = ((got.float() - full.float()).abs().max() / full.float().abs().max()).item()
            msgs.append(("encprocdec", str(dt), err, err <= (1e-5 if dt == torch.float32 else 2e-2) and got.shape == full.shape))
        ret[rank] = msgs
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[rank] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


def test_sharded_processors_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(world, os.path.join(d, "rdv"), ret), nprocs=world, join=True)
        for r in range(world):
            assert isinstance(ret.get(r), list), ret.get(r)
            for kind, dt, err, ok in ret[r]:
                assert ok, f"rank {r} {kind} {dt}: sharded vs single rel err {err:.3e}"
