"""ConditionalLayerNorm (reference layers/normalization.py:34-94): the module mirrors the reference parameter names,
the oracle reproduces a golden of the unmodified reference GraphTransformerProcessor run with ConditionalLayerNorm kernels and ``cond=``,
and the CUDA path (one conditional-LayerNorm kernel in front of each GEMM) matches that golden."""
import pytest
import torch

from anemoi_core_b200.layers import GraphTransformerProcessor
from anemoi_core_b200.layers.normalization import ConditionalLayerNorm
from oracle import restatement as R


def _build(g):
    lk = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": g["condition_shape"],
                        "zero_init": False}}  # fmt: skip
    m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **g["cfg"]).eval()
    assert sorted(m.state_dict().keys()) == sorted(g["sd"].keys())  # ...layer_norm_attention.scale.weight / .bias.weight ...
    m.load_state_dict(g["sd"], strict=True)
    return m


def test_oracle_matches_reference_golden(golden):
    g = golden("gt_processor_condln")
    y = R.gt_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"], g["cfg"]["num_heads"], cond=g["cond"])
    torch.testing.assert_close(y, g["y"], atol=2e-5, rtol=1e-5)


def test_module_mirrors_reference(golden):
    m = _build(golden("gt_processor_condln"))
    ln = m.proc[0].layer_norm_attention
    assert isinstance(ln, ConditionalLayerNorm) and isinstance(m.proc[1].layer_norm_mlp_dst, ConditionalLayerNorm)
    assert ln.scale.weight.shape == (64, 16) and not any(True for _ in ln.norm.parameters())
    z = ConditionalLayerNorm(8, condition_shape=4)  # zero_init: starts as a plain LayerNorm
    assert all(float(p.abs().sum()) == 0.0 for p in z.parameters())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        z(torch.randn(3, 8), torch.randn(3, 4))


@pytest.mark.gpu
def test_gt_processor_with_conditional_layer_norm(golden):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    g = golden("gt_processor_condln")
    m = _build(g).cuda()
    args = (g["x"].cuda(), 1, GraphShardInfo(), g["edge_attr"].cuda(), g["edge_index"].cuda())
    with torch.no_grad():
        y32 = m(*args, cond=g["cond"].cuda())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(*args, cond=g["cond"].cuda())
        with pytest.raises(ValueError, match="conditioning"):
            m(*args)
    ref = g["y"]
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C,Dc", [(1000, 512, 16), (33, 100, 7), (5, 8, 32), (40962, 1024, 16)])
def test_cond_layer_norm_kernel(dt, M, C, Dc):
    ln = ConditionalLayerNorm(C, condition_shape=Dc, zero_init=False)
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 1.5 + 0.3).to(dt)
    cond = torch.randn(M, Dc, generator=g)
    xf = x.float()
    ref = torch.nn.functional.layer_norm(xf, (C,)) * (1 + ln.scale(cond)) + ln.bias(cond)
    with torch.no_grad():
        y = ln.cuda()(x.cuda(), cond.cuda())
    assert y.dtype == dt
    tol = 2e-5 if dt == torch.float32 else 2**-7
    assert ((y.float().cpu() - ref.detach()).abs() <= tol * ref.detach().abs().clamp_min(1.0)).all()


def _build_mapper(g):
    from anemoi_core_b200.layers import GraphTransformerForwardMapper

    lk = {"LayerNorm": {"_target_": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm", "condition_shape": g["condition_shape"],
                        "zero_init": False}}  # fmt: skip
    m = GraphTransformerForwardMapper(num_chunks=1, mlp_hidden_ratio=4, layer_kernels=lk, **g["mapper"]["cfg"]).eval()
    assert sorted(m.state_dict().keys()) == sorted(g["mapper"]["sd"].keys())
    m.load_state_dict(g["mapper"]["sd"], strict=True)
    return m


def test_oracle_mapper_with_cond(golden):
    g = golden("gt_processor_condln")
    mp = g["mapper"]
    _, yd = R.gt_forward_mapper(mp["sd"], mp["x_src"], mp["x_dst"], mp["edge_attr"], mp["edge_index"], mp["cfg"]["num_heads"],
                                cond=(mp["cond_src"], mp["cond_dst"]))  # fmt: skip
    torch.testing.assert_close(yd, mp["y_dst"], atol=2e-5, rtol=1e-5)
    _build_mapper(g)  # constructs with the reference state_dict keys on CPU


@pytest.mark.gpu
def test_gt_forward_mapper_with_cond(golden):
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo

    g = golden("gt_processor_condln")
    mp = g["mapper"]
    m = _build_mapper(g).cuda()
    args = ((mp["x_src"].cuda(), mp["x_dst"].cuda()), 1, BipartiteGraphShardInfo(), mp["edge_attr"].cuda(), mp["edge_index"].cuda())
    cond = (mp["cond_src"].cuda(), mp["cond_dst"].cuda())
    with torch.no_grad():
        _, y32 = m(*args, cond=cond)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, y16 = m(*args, cond=cond)
    ref = mp["y_dst"]
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2
