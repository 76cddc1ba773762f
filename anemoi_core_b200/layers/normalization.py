"""LayerNorm containers (reference: layers/normalization.py:19-31)."""

from __future__ import annotations

import torch
from torch import Tensor
from torch import nn

from .. import ops


class AutocastLayerNorm(nn.LayerNorm):
    """LayerNorm that returns the input dtype (the reference casts back with ``type_as``).  Used standalone for
    q/k normalisation; the forward is the fused sm_100a kernel."""

    def forward(self, x: Tensor) -> Tensor:
        shape = x.shape
        y = ops.layer_norm(x.reshape(-1, shape[-1]), self.weight, self.bias, self.eps, out_dtype=x.dtype)
        return y.reshape(shape)


def ln_params(ln: nn.Module) -> tuple[Tensor | None, Tensor | None, float]:
    """(weight, bias, eps) of a LayerNorm-like container."""
    if isinstance(ln, nn.Identity):
        raise TypeError("expected a LayerNorm container")
    return getattr(ln, "weight", None), getattr(ln, "bias", None), float(getattr(ln, "eps", 1e-5))


def _check_plain_layernorm(ln: nn.Module) -> None:
    if not isinstance(ln, torch.nn.LayerNorm):
        raise NotImplementedError(
            f"{type(ln).__name__}: only torch.nn.LayerNorm-like kernels (weight, bias, eps) are implemented; "
            "ConditionalLayerNorm is outside the forward hot path (SURVEY.md §8f rank 4)"
        )
