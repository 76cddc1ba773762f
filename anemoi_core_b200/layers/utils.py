"""Layer-kernel factory and small helpers (reference: layers/utils.py:25-52, 87-142).

``layer_kernels`` keeps the reference's plugin shape — ``{"Linear": {"_target_": "torch.nn.Linear", ...}, ...}`` —
but here the resolved classes are *parameter containers*: the forward of every block runs fused sm_100a kernels
that read ``.weight`` / ``.bias`` directly, so a custom kernel class only has to expose those attributes.
Resolution is by plain ``importlib`` (no Hydra needed); a Hydra DictConfig / DotDict works too (mapping access).
"""

from __future__ import annotations

import functools
import importlib
import math
from typing import Any
from typing import Mapping
from typing import Optional


class LayerKernels(dict):
    """dict with attribute access (the reference uses anemoi.utils.config.DotDict)."""

    def __getattr__(self, name: str):
        try:
            return self[name]
        except KeyError:  # AttributeError, not KeyError: copy.deepcopy / pickle probe __deepcopy__ / __getstate__ with getattr(obj, name, None)
            raise AttributeError(name) from None


DEFAULT_KERNELS = {
    "Linear": {"_target_": "torch.nn.Linear"},
    "LayerNorm": {"_target_": "torch.nn.LayerNorm"},
    "Activation": {"_target_": "torch.nn.GELU"},
    "QueryNorm": {"_target_": "anemoi_core_b200.layers.normalization.AutocastLayerNorm", "bias": False},
    "KeyNorm": {"_target_": "anemoi_core_b200.layers.normalization.AutocastLayerNorm", "bias": False},
}

# reference dotted paths are accepted and mapped onto the in-package container classes
_ALIASES = {
    "anemoi.models.layers.normalization.AutocastLayerNorm": "anemoi_core_b200.layers.normalization.AutocastLayerNorm",
    "anemoi.models.layers.normalization.ConditionalLayerNorm": "anemoi_core_b200.layers.normalization.ConditionalLayerNorm",
}


def _resolve(path: str) -> Any:
    path = _ALIASES.get(path, path)
    mod, _, name = path.rpartition(".")
    try:
        return getattr(importlib.import_module(mod), name)
    except (ImportError, AttributeError) as e:
        raise ImportError(f"layer_kernels: cannot import {path!r}: {e}") from e


def load_layer_kernels(kernel_config: Optional[Mapping] = None, instance: bool = True) -> LayerKernels:
    """Factories for Linear / LayerNorm / Activation / QueryNorm / KeyNorm; missing entries default to torch.nn
    (layers/utils.py:87-142).  An already-resolved LayerKernels is returned unchanged."""
    if isinstance(kernel_config, LayerKernels):
        return kernel_config
    merged = {**DEFAULT_KERNELS, **dict(kernel_config or {})}
    out = LayerKernels()
    for name, entry in merged.items():
        if not instance:
            out[name] = entry
            continue
        if callable(entry):
            out[name] = entry
            continue
        entry = dict(entry)
        target = entry.pop("_target_")
        entry.pop("_partial_", None)
        cls = _resolve(target) if isinstance(target, str) else target
        out[name] = functools.partial(cls, **entry) if entry else cls
    return out


def compute_mlp_hidden_dim(num_channels: int, mlp_hidden_ratio: float) -> int:
    """int(num_channels * ratio + 0.5), validated like layers/utils.py:25-52."""
    if not math.isfinite(mlp_hidden_ratio):
        raise ValueError(f"`mlp_hidden_ratio` must be finite, got {mlp_hidden_ratio}.")
    if mlp_hidden_ratio <= 0:
        raise ValueError(f"`mlp_hidden_ratio` must be > 0, got {mlp_hidden_ratio}.")
    hidden = int(num_channels * mlp_hidden_ratio + 0.5)
    if hidden <= 0:
        raise ValueError(f"Computed hidden_dim must be > 0, got {hidden}.")
    return hidden
