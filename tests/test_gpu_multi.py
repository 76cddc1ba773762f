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
        # ---- (round 2) the same step with sharded inputs / outputs, every exchange a peer-memory kernel, captured as ONE CUDA graph ----
        from anemoi_core_b200.distributed import peer

        msgs.append(("peer_memory_available", "-", 0.0, peer.available(group, torch.device("cuda", rank))))
        g0, m0 = sum(gsz[:rank]), sum(sizes[:rank])
        xg_l, xm_l = xg[g0 : g0 + gsz[rank]].contiguous(), xm[m0 : m0 + sizes[rank]].contiguous()
        kw = dict(model_comm_group=group, mesh_shards=sizes, grid_shards=gsz, keep_output_sharded=True, inputs_sharded=True)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            full = model(xg, xm, gd)
            got = model(xg_l, xm_l, gd, **kw)
            err = ((got.float() - full[g0 : g0 + gsz[rank]].float()).abs().max() / full.float().abs().max()).item()
            msgs.append(("encprocdec_sharded_io", "bf16", err, err <= 2e-2 and got.shape[0] == gsz[rank]))
            if peer.available(group):
                replay = model.capture(xg_l, xm_l, gd, **kw)
                for it in range(3):  # replays keep the device-side exchange counters in step across ranks
                    rep = replay().clone()
                    err = ((rep.float() - got.float()).abs().max() / got.float().abs().max().clamp_min(1e-30)).item()
                    msgs.append((f"whole_graph_replay_{it}", "bf16", err, err <= 1e-6))
                # new inputs through the static buffers
                rep = replay(xg_l * 0.5, xm_l).clone()
                ref2 = model(xg_l * 0.5, xm_l, gd, **kw)
                err = ((rep.float() - ref2.float()).abs().max() / ref2.float().abs().max()).item()
                msgs.append(("whole_graph_replay_new_input", "bf16", err, err <= 1e-6))
        # ---- NCCL all-to-all fallback of the halo exchange (ANEMOI_B200_PEER=0) must give the same rows ----
        from anemoi_core_b200.distributed import halo

        peer._STATE[id(group)] = False
        halo._PLANS.clear()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            got2 = model(xg_l, xm_l, gd, **kw)
        err = ((got2.float() - got.float()).abs().max() / got.float().abs().max()).item()
        msgs.append(("nccl_fallback_vs_peer", "bf16", err, err <= 1e-6))
        del peer._STATE[id(group)]
        halo._PLANS.clear()
        # ---- heads (Ulysses) strategy and the GNN halo form, first time on NCCL (Gloo-only in round 1) ----
        from anemoi_core_b200.layers import processor as P

        torch.manual_seed(0)
        mh = GraphTransformerProcessor(num_layers=2, num_channels=256, num_chunks=1, num_heads=8, mlp_hidden_ratio=4, edge_dim=gr["edge_dim"],
                                       shard_strategy="heads").cuda().eval()  # fmt: skip
        x = torch.randn(n, 256, generator=torch.Generator().manual_seed(1)).cuda()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            full = mh(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
            local = mh(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
            gath = gather_rows(local, sizes, group)
        err = ((gath.float() - full.float()).abs().max() / full.float().abs().max()).item()
        msgs.append(("gt_heads_strategy", "bf16", err, err <= 2e-2))
        # heads strategy in TRAINING (round 2): one fp32 step of the sharded processor gives the single-GPU gradients (AllToAllFn)
        torch.manual_seed(3)
        mh = GraphTransformerProcessor(num_layers=2, num_channels=64, num_chunks=1, num_heads=4, mlp_hidden_ratio=4, edge_dim=gr["edge_dim"],
                                       shard_strategy="heads", qk_norm=True).cuda().train()  # fmt: skip
        x0 = torch.randn(n, 64, generator=torch.Generator().manual_seed(4)).cuda()
        wt = torch.randn(n, 64, generator=torch.Generator().manual_seed(5)).cuda()
        xf = x0.clone().requires_grad_()
        (mh(xf, 1, GraphShardInfo(nodes=[n]), ea, ei) * wt).sum().backward()
        ref_p = {n_: p.grad.clone() for n_, p in mh.named_parameters()}
        ref_x = xf.grad.clone()
        mh.zero_grad()
        xs_ = shard_rows(x0, sizes, group).contiguous().clone().requires_grad_()
        y_l = mh(xs_, 1, GraphShardInfo(nodes=sizes), ea, ei, group)
        (y_l * shard_rows(wt, sizes, group)).sum().backward()
        rx = shard_rows(ref_x, sizes, group)
        err = ((xs_.grad - rx).abs().max() / rx.abs().max()).item()
        msgs.append(("train_heads_dx", "fp32", err, err <= 1e-4))
        worst, big = 0.0, max(g_.abs().max().item() for g_ in ref_p.values())
        for n_, p in mh.named_parameters():
            gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
            dist.all_reduce(gp)
            worst = max(worst, ((gp - ref_p[n_]).abs().max() / max(ref_p[n_].abs().max().item(), 1e-3 * big)).item())
        msgs.append(("train_heads_dparams", "fp32", worst, worst <= 2e-4))
        P.GNN_HALO = True
        torch.manual_seed(0)
        mg = GNNProcessor(num_channels=128, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=gr["edge_dim"]).cuda().eval()
        x = torch.randn(n, 128, generator=torch.Generator().manual_seed(1)).cuda()
        with torch.no_grad():
            full = mg(x, 1, GraphShardInfo(nodes=[n]), ea, ei)
            local = mg(shard_rows(x, sizes, group).contiguous(), 1, GraphShardInfo(nodes=sizes), ea, ei, group)
            gath = gather_rows(local, sizes, group)
        err = ((gath - full).abs().max() / full.abs().max()).item()
        msgs.append(("gnn_halo_form", "fp32", err, err <= 1e-5))
        P.GNN_HALO = False
        # ---- (round 2) TRAINING on the model-parallel group: the autograd halves of the halo exchange / all-gather (reference
        # distributed/graph.py:227-500).  One fp32 step of the sharded model must give the single-GPU gradients: parameter gradients are
        # per-rank partial sums (all-reduced here, as the trainer does over the model group), input gradients are the local slices. ----
        wgt = torch.randn(gr["n_grid"], 9, generator=torch.Generator().manual_seed(21)).cuda()
        for kind in ("graphtransformer", "gnn"):
            torch.manual_seed(6)
            mt = EncProcDec(kind, in_grid=20, in_mesh=12, out_grid=9, num_channels=64, num_layers=2, edge_dim=gr["edge_dim"], num_heads=4).cuda().train()
            xg_f = xg.detach().clone().requires_grad_()
            (mt(xg_f, xm, gd) * wgt).sum().backward()
            ref_p = {n_: p.grad.clone() for n_, p in mt.named_parameters()}
            ref_x = xg_f.grad.clone()
            mt.zero_grad()
            if kind == "graphtransformer":
                xg_s = xg[g0 : g0 + gsz[rank]].detach().clone().requires_grad_()
                y_l = mt(xg_s, xm[m0 : m0 + sizes[rank]].contiguous(), gd, **kw)
                (y_l * wgt[g0 : g0 + gsz[rank]]).sum().backward()
                gx, rx = xg_s.grad, ref_x[g0 : g0 + gsz[rank]]
            else:
                xg_s = xg.detach().clone().requires_grad_()
                y_all = mt(xg_s, xm, gd, group, sizes, gsz)  # replicated in / gathered out: every rank holds the whole output
                (y_all * wgt).sum().backward()
                gx = xg_s.grad.clone()
                dist.all_reduce(gx)
                gx, rx = gx / world, ref_x  # every rank back-propagated the same (whole) loss
            err = ((gx - rx).abs().max() / rx.abs().max()).item()
            msgs.append((f"train_{kind}_dx", "fp32", err, err <= 1e-4))
            worst = 0.0
            big = max(g_.abs().max().item() for g_ in ref_p.values())
            for n_, p in mt.named_parameters():
                gp = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
                dist.all_reduce(gp)
                if kind == "gnn":
                    gp = gp / world
                worst = max(worst, ((gp - ref_p[n_]).abs().max() / max(ref_p[n_].abs().max().item(), 1e-3 * big)).item())
            msgs.append((f"train_{kind}_dparams", "fp32", worst, worst <= 2e-4))
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
            print(f"rank {r}:", [(k, d_, f"{e:.2e}") for k, d_, e, _ in ret[r]])
