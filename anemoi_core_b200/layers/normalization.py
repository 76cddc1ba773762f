"""LayerNorm containers (reference: layers/normalization.py:19-94)."""

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


class ConditionalLayerNorm(nn.Module):
    """``LN(x) * (1 + scale(cond)) + bias(cond)`` (normalization.py:34-94), same parameter names (``scale.weight|bias``, ``bias.weight|bias``;
    the LayerNorm itself has no affine).  One kernel: the per-row scale / bias are formed inside it, never materialised."""

    def __init__(self, normalized_shape: int, condition_shape: int = 16, zero_init: bool = True, autocast: bool = True) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(normalized_shape, elementwise_affine=False)
        self.scale = nn.Linear(condition_shape, normalized_shape)
        self.bias = nn.Linear(condition_shape, normalized_shape)
        self.autocast = autocast
        if zero_init:
            for p in (self.scale.weight, self.scale.bias, self.bias.weight, self.bias.bias):
                nn.init.zeros_(p)

    @property
    def eps(self) -> float:
        return self.norm.eps

    def run(self, x: Tensor, cond: Tensor, out_dtype: torch.dtype) -> Tensor:
        if cond is None:
            raise ValueError("ConditionalLayerNorm needs the conditioning tensor (cond=...)")
        return ops.cond_layer_norm(x, cond, self.scale.weight.detach(), self.scale.bias.detach(), self.bias.weight.detach(), self.bias.bias.detach(),
                                   self.eps, out_dtype=out_dtype)  # fmt: skip

    def forward(self, x: Tensor, cond: Tensor) -> Tensor:
        shape = x.shape
        y = self.run(x.reshape(-1, shape[-1]), cond.reshape(-1, cond.shape[-1]), x.dtype if self.autocast else torch.float32)
        return y.reshape(shape)


def ln_params(ln: nn.Module) -> tuple[Tensor | None, Tensor | None, float]:
    """(weight, bias, eps) of a LayerNorm-like container."""
    if isinstance(ln, nn.Identity):
        raise TypeError("expected a LayerNorm container")
    return getattr(ln, "weight", None), getattr(ln, "bias", None), float(getattr(ln, "eps", 1e-5))


def _check_plain_layernorm(ln: nn.Module) -> None:
    if not isinstance(ln, torch.nn.LayerNorm):
        raise NotImplementedError(
            f"{type(ln).__name__}: only torch.nn.LayerNorm-like kernels (weight, bias, eps) and ConditionalLayerNorm are implemented"
        )
