"""Constructor options outside the default configs: ``edge_pre_mlp=True``, ``attn_channels != num_channels``,
``mlp_extra_layers > 0`` — goldens of the unmodified reference processors; oracle on CPU, CUDA path on the GPU (fp32 1e-4, bf16 2e-2)."""
import pytest
import torch

from anemoi_core_b200.layers import GNNProcessor
from anemoi_core_b200.layers import GraphTransformerProcessor
from oracle import restatement as R

GT = ["gt_processor_edge_pre_mlp", "gt_processor_attn_channels"]


@pytest.mark.parametrize("name", GT)
def test_oracle_gt_options(golden, name):
    g = golden(name)
    y = R.gt_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"], g["cfg"]["num_heads"])
    torch.testing.assert_close(y, g["y"], atol=2e-5, rtol=1e-5)


def test_oracle_gnn_extra_layers(golden):
    g = golden("gnn_processor_extra_layers")
    y = R.gnn_processor(g["sd"], g["x"], g["edge_attr"], g["edge_index"], g["cfg"]["num_layers"])
    torch.testing.assert_close(y, g["y"], atol=2e-5, rtol=1e-5)


def _run(m, g, shard):
    args = (g["x"].cuda(), 1, shard, g["edge_attr"].cuda(), g["edge_index"].cuda())
    with torch.no_grad():
        y32 = m(*args)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(*args)
    ref = g["y"]
    assert (y32.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert ((y16.float().cpu() - ref).norm() / ref.norm()) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name", GT)
def test_gt_options_cuda(golden, name):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    g = golden(name)
    m = GraphTransformerProcessor(num_chunks=1, mlp_hidden_ratio=4, **g["cfg"]).eval()
    assert sorted(m.state_dict().keys()) == sorted(g["sd"].keys())
    m.load_state_dict(g["sd"], strict=True)
    _run(m.cuda(), g, GraphShardInfo())


@pytest.mark.gpu
def test_gnn_extra_layers_cuda(golden):
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    g = golden("gnn_processor_extra_layers")
    m = GNNProcessor(num_chunks=1, **g["cfg"]).eval()
    assert sorted(m.state_dict().keys()) == sorted(g["sd"].keys())
    m.load_state_dict(g["sd"], strict=True)
    _run(m.cuda(), g, GraphShardInfo(nodes=[g["x"].shape[0]]))
