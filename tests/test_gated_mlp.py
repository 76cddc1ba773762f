"""Gated feed-forward variants (mlp_implementation = glu / swiglu / geglu / reglu; reference layers/mlp.py:27-94):
oracle vs goldens of the unmodified reference processors (CPU), reference state_dict keys, and the CUDA path (one GEMM on the
concatenated gate | value weights + ``glu_combine``) vs the goldens in fp32 and bf16."""
import pytest
import torch

from anemoi_core_b200.layers import GNNProcessor
from anemoi_core_b200.layers import GraphTransformerProcessor
from anemoi_core_b200.layers.mlp import MLP
from anemoi_core_b200.layers.mlp import GatedMLPLayer
from oracle import restatement as R

KINDS = ["glu", "swiglu", "geglu", "reglu"]


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_gt_processor_gated(golden, kind):
    g = golden(f"gt_processor_{kind}")
    assert g["cfg"]["mlp_implementation"] == kind
    with R.gated_mlp(kind):
        y = R.gt_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"], g["cfg"]["num_heads"])
    torch.testing.assert_close(y, g["y"], atol=2e-5, rtol=1e-5)
    if kind != "swiglu":  # the gating really is part of the function: the wrong one does not reproduce the golden
        with R.gated_mlp("swiglu"):
            y_wrong = R.gt_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"], g["cfg"]["num_heads"])
        assert (y_wrong - g["y"]).abs().max() > 1e-3


def test_oracle_gnn_processor_gated(golden):
    g = golden("gnn_processor_swiglu")
    with R.gated_mlp("swiglu"):
        y = R.gnn_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"])
    torch.testing.assert_close(y, g["y"], atol=2e-5, rtol=1e-5)


def _build(g, cls):
    cfg = dict(g["cfg"])
    if cls is GraphTransformerProcessor:
        cfg.update(num_chunks=1, mlp_hidden_ratio=4)
    else:
        cfg.update(num_chunks=1, mlp_extra_layers=0)
    m = cls(**cfg).eval()
    assert sorted(m.state_dict().keys()) == sorted(g["sd"].keys())  # mlp.0.gate_proj / mlp.0.value_proj / mlp.1 ...
    m.load_state_dict(g["sd"], strict=True)
    return m


def test_gated_modules_mirror_the_reference(golden):
    m = _build(golden("gt_processor_geglu"), GraphTransformerProcessor)
    first = m.proc[0].node_dst_mlp.mlp[0]
    assert isinstance(first, GatedMLPLayer) and isinstance(first.gating, torch.nn.GELU) and len(m.proc[0].node_dst_mlp.mlp) == 2
    _build(golden("gnn_processor_swiglu"), GNNProcessor)
    with pytest.raises(ValueError):
        MLP(8, 8, 8, mlp_implementation="nope")


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gt_processor_gated_cuda(golden, kind):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    g = golden(f"gt_processor_{kind}")
    m = _build(g, GraphTransformerProcessor).cuda()
    args = (g["x"].cuda(), 1, GraphShardInfo(), g["edge_attr"].cuda(), g["edge_index"].cuda())
    with torch.no_grad():
        y32 = m(*args)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(*args)
    ref = g["y"]
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2


@pytest.mark.gpu
def test_gnn_processor_gated_cuda(golden):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    g = golden("gnn_processor_swiglu")
    m = _build(g, GNNProcessor).cuda()
    n = g["x"].shape[0]
    args = (g["x"].cuda(), 1, GraphShardInfo(nodes=[n]), g["edge_attr"].cuda(), g["edge_index"].cuda())
    with torch.no_grad():
        y32 = m(*args)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(*args)
    ref = g["y"]
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,H", [(1000, 256), (33, 20), (1, 1), (40962, 2048)])
def test_glu_combine_kernel(kind, dt, M, H):
    from anemoi_core_b200 import ops

    gv = torch.randn(M, 2 * H, generator=torch.Generator().manual_seed(M + H)).to(dt)
    gate = {"glu": torch.sigmoid, "swiglu": torch.nn.functional.silu, "geglu": torch.nn.functional.gelu, "reglu": torch.relu}[kind]
    ref = gate(gv[:, :H].float()) * gv[:, H:].float()
    y = ops.glu_combine(gv.cuda(), kind)
    assert y.shape == (M, H) and y.dtype == dt
    tol = 2e-6 if dt == torch.float32 else 2**-8
    assert ((y.float().cpu() - ref).abs() <= tol * ref.abs().clamp_min(1.0)).all()
