"""The reference's `_target_` strings resolve to the B200 classes after `overlay.install()` (with `config_validation: True`
pydantic only accepts the reference's literal paths, schemas/processor.py:29,38, encoder.py:25,30, decoder.py:25,30)."""
import importlib
import sys

import pytest


def hydra_like_get_class(path: str):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


# training/src/anemoi/training/config/model/graphtransformer.yaml:27,44,60 and gnn.yaml:22,35,47
YAML_TARGETS = [
    "anemoi.models.layers.processor.GraphTransformerProcessor",
    "anemoi.models.layers.mapper.GraphTransformerForwardMapper",
    "anemoi.models.layers.mapper.GraphTransformerBackwardMapper",
    "anemoi.models.layers.processor.GNNProcessor",
    "anemoi.models.layers.mapper.GNNForwardMapper",
    "anemoi.models.layers.mapper.GNNBackwardMapper",
]


@pytest.fixture()
def overlay():
    from anemoi_core_b200 import overlay as ov

    before = {k: v for k, v in sys.modules.items() if k == "anemoi" or k.startswith("anemoi.")}
    ov.install()
    yield ov
    ov.uninstall()
    for k in [k for k in sys.modules if (k == "anemoi" or k.startswith("anemoi.")) and k not in before]:
        del sys.modules[k]


def test_reference_targets_resolve_to_b200_classes(overlay):
    import anemoi_core_b200.layers as L

    assert set(YAML_TARGETS) <= set(overlay.targets())
    for path in overlay.targets():
        cls = hydra_like_get_class(path)
        assert cls is getattr(L, path.rsplit(".", 1)[1]) and cls.__module__.startswith("anemoi_core_b200.")


def test_instantiate_from_the_reference_yaml_processor_block(overlay):
    """graphtransformer.yaml:26-41 (processor block) as Hydra would pass it: `_target_` + kwargs, incl. keys this forward ignores."""
    cfg = {
        "_target_": "anemoi.models.layers.processor.GraphTransformerProcessor",
        "trainable_size": 8, "sub_graph_edge_attributes": ["edge_length", "edge_dirs"], "num_layers": 2, "num_chunks": 2, "num_heads": 4,
        "mlp_hidden_ratio": 4, "qk_norm": False, "cpu_offload": False, "gradient_checkpointing": True, "graph_attention_backend": "triton",
        "edge_pre_mlp": False,
        "layer_kernels": {"LayerNorm": {"_target_": "anemoi.models.layers.normalization.AutocastLayerNorm"}},
    }  # fmt: skip
    kwargs = {k: v for k, v in cfg.items() if k != "_target_"}
    m = hydra_like_get_class(cfg["_target_"])(num_channels=64, edge_dim=11, **kwargs)
    assert type(m).__module__ == "anemoi_core_b200.layers.processor" and len(m.proc) == 2
    # same state_dict keys as the reference block: a reference checkpoint loads with strict=True
    keys = set(m.state_dict())
    assert {"proc.0.lin_query.weight", "proc.0.lin_edge.bias", "proc.1.node_dst_mlp.mlp.2.weight", "proc.0.layer_norm_attention.weight"} <= keys


def test_uninstall_restores(overlay):
    overlay.uninstall()
    assert not any(k == "anemoi" or k.startswith("anemoi.") for k in sys.modules if getattr(sys.modules[k], "__anemoi_b200_stub__", False) is True)
    overlay.install()
