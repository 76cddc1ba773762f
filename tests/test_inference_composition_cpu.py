"""The INFERENCE composition of the drop-in modules (layers/block.py, processor.py, mapper.py: LayerNorm folded into the consuming GEMMs, row
statistics handed from producer to consumer, packed q|k|v|self|qw weights, lin_edge folded into attention, the gather-add first edge GEMM, the
one-kernel GraphConv routing) against the golden outputs of the UNMODIFIED reference modules, on CPU: ``tests/_cpu_ops.py`` stands in for the
CUDA entry points with plain fp32 PyTorch statements of each fused op, so what is checked here is the HOST side — the algebra of the folds and
the call sequence — against the reference's own numbers (fp32 bar of SURVEY.md §8d: 1e-4 of the output's scale).  The kernels themselves are
pinned by ``-m gpu`` on the same fixtures (tests/test_gpu_parity.py, test_constructor_options.py, test_gated_mlp.py, test_cond_layer_norm.py).
Runs in a spawned process because the stand-ins replace ``anemoi_core_b200.ops`` globally."""
import os
import sys
import tempfile

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _g(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def _err(y, ref):
    return ((y.float() - ref).abs().max() / ref.abs().max()).item()


def _run_all():
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNBackwardMapper
    from anemoi_core_b200.layers import GNNForwardMapper
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerBackwardMapper
    from anemoi_core_b200.layers import GraphTransformerForwardMapper
    from anemoi_core_b200.layers import GraphTransformerProcessor

    out = []
    with torch.no_grad():
        for name in ("gnn_processor_small", "gnn_processor_cfg1", "gnn_processor_extra_layers", "gnn_processor_swiglu"):
            g = _g(name)
            cfg = dict(g["cfg"])
            cfg.setdefault("mlp_extra_layers", 0)
            m = GNNProcessor(num_chunks=1, **cfg).eval()
            m.load_state_dict(g["sd"], strict=True)
            out.append((name, _err(m(g["x"], 1, GraphShardInfo(nodes=[g["x"].shape[0]]), g["edge_attr"], g["edge_index"]), g["y"])))
        for name in ("gt_processor_small", "gt_processor_qknorm", "gt_processor_unsorted", "gt_processor_edge_pre_mlp", "gt_processor_attn_channels",
                     "gt_processor_glu", "gt_processor_swiglu", "gt_processor_geglu", "gt_processor_reglu"):  # fmt: skip
            g = _g(name)
            m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, **g["cfg"]).eval()
            m.load_state_dict(g["sd"], strict=True)
            y = m(g["x"], 1, GraphShardInfo(), g["edge_attr"], g["edge_index"], edges_are_dst_sorted=g.get("sorted", True))
            out.append((name, _err(y, g["y"])))
        sh = BipartiteGraphShardInfo()
        for name, cls, kw in (("gnn_forward_mapper", GNNForwardMapper, dict(num_chunks=1, mlp_extra_layers=0)),
                              ("gnn_backward_mapper", GNNBackwardMapper, dict(num_chunks=1, mlp_extra_layers=0)),
                              ("gt_forward_mapper_chunks1", GraphTransformerForwardMapper, dict(mlp_hidden_ratio=4)),
                              ("gt_forward_mapper_chunks4", GraphTransformerForwardMapper, dict(mlp_hidden_ratio=4)),
                              ("gt_backward_mapper", GraphTransformerBackwardMapper, dict(mlp_hidden_ratio=4))):  # fmt: skip
            g = _g(name)
            m = cls(**g["cfg"], **kw).eval()
            m.load_state_dict(g["sd"], strict=True)
            y = m((g["x_src"], g["x_dst"]), 1, sh, g["edge_attr"], g["edge_index"])
            if isinstance(y, tuple):
                if "y_src" in g and cls is GNNForwardMapper:
                    out.append((name + ".src", _err(y[0], g["y_src"])))
                out.append((name + ".dst", _err(y[1], g["y_dst"])))
            else:
                out.append((name, _err(y, g["y"])))
        # ConditionalLayerNorm kernels with cond= (processor and forward mapper; stand-in: cond_layer_norm)
        g = _g("gt_processor_condln")
        lk = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": g["condition_shape"],
                            "zero_init": False}}  # fmt: skip
        m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **g["cfg"]).eval()
        m.load_state_dict(g["sd"], strict=True)
        out.append(("gt_processor_condln", _err(m(g["x"], 1, GraphShardInfo(), g["edge_attr"], g["edge_index"], cond=g["cond"]), g["y"])))
        mp_ = g["mapper"]
        m = GraphTransformerForwardMapper(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **mp_["cfg"]).eval()
        m.load_state_dict(mp_["sd"], strict=True)
        _, yd = m((mp_["x_src"], mp_["x_dst"]), 1, sh, mp_["edge_attr"], mp_["edge_index"], cond=(mp_["cond_src"], mp_["cond_dst"]))
        out.append(("gt_forward_mapper_condln", _err(yd, mp_["y_dst"])))
        # the whole AnemoiModelEncProcDec forward (graph providers, assemble_input / assemble_output, latent skip, boundings): both model kinds
        from test_model_glue import build_model

        fx = _g("model_forward")
        for kind in ("graphtransformer", "gnn"):
            m = build_model(fx, kind)
            m.load_state_dict(fx["cases"][kind]["sd"], strict=True)
            y = m({"data": fx["x"]})["data"]
            assert (y[..., fx["bound_vars"]] >= 0).all()
            out.append((f"model_forward_{kind}", _err(y, fx["cases"][kind]["y"])))
    return out


def _worker(rank, init_file, ret):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=0, world_size=1)
    try:
        import _cpu_ops

        _cpu_ops.install()
        torch.set_num_threads(4)
        ret[0] = _run_all()
    except Exception as e:  # noqa: BLE001
        import traceback

        ret[0] = f"{type(e).__name__}: {e}\n{traceback.format_exc()}"
    finally:
        dist.destroy_process_group()


def test_inference_composition_matches_reference_goldens_cpu():
    with tempfile.TemporaryDirectory() as d:
        ret = mp.Manager().dict()
        mp.spawn(_worker, args=(os.path.join(d, "rdv"), ret), nprocs=1, join=True)
        assert isinstance(ret.get(0), list), ret.get(0)
        assert len(ret[0]) == 23
        for name, err in ret[0]:
            assert err <= 1e-4, f"{name}: max|ours - reference| / max|reference| = {err:.3e}"
        print([(n, f"{e:.1e}") for n, e in ret[0]])
