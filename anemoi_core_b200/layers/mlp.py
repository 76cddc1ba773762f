"""MLP container + fused forward (reference: layers/mlp.py:97-179).

Same structure and parameter names as the reference — ``mlp`` = Sequential(Linear, act, [Linear, act]*k, Linear[, act]),
optional ``layer_norm`` — so reference ``state_dict``s load unchanged.  The forward issues one fused
GEMM(+bias+GELU) kernel per Linear and one LayerNorm(+residual) kernel.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch import nn

from . import _functional as Fn
from .utils import load_layer_kernels


def _is_gelu(m: nn.Module) -> bool:
    return isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none"


class MLP(nn.Module):
    def __init__(
        self,
        in_features: int,
        hidden_dim: int,
        out_features: int,
        layer_kernels=None,
        n_extra_layers: int = 0,
        final_activation: bool = False,
        layer_norm: bool = True,
        mlp_implementation: str = "mlp",
    ) -> None:
        super().__init__()
        if n_extra_layers < 0:
            raise ValueError(f"`n_extra_layers` must be >= 0, got {n_extra_layers}.")
        if mlp_implementation != "mlp":
            raise NotImplementedError(
                f"mlp_implementation={mlp_implementation!r}: gated variants (glu/swiglu/geglu/reglu) are not part of the "
                "implemented hot path yet (SURVEY.md §8f rank 4)"
            )
        k = load_layer_kernels(layer_kernels)
        layers: list[nn.Module] = [k.Linear(in_features, hidden_dim), k.Activation()]
        for _ in range(n_extra_layers):
            layers += [k.Linear(hidden_dim, hidden_dim), k.Activation()]
        layers.append(k.Linear(hidden_dim, out_features))
        if final_activation:
            layers.append(k.Activation())
        for m in layers:
            if not hasattr(m, "weight") and not _is_gelu(m):
                raise NotImplementedError(f"Activation {type(m).__name__}: only exact (erf) torch.nn.GELU is fused into the GEMM epilogue")
        self.mlp = nn.Sequential(*layers)
        self.layer_norm = k.LayerNorm(normalized_shape=out_features) if layer_norm else None
        self._pack = Fn.WeightPack()

    def run(
        self,
        x: Tensor,
        dt: torch.dtype,
        residual: Optional[Tensor] = None,
        first_gathers: Optional[tuple] = None,
        first_cols: Optional[slice] = None,
        out: Optional[Tensor] = None,
        pre_ln: Optional[nn.Module] = None,
        want_stats: bool = False,
    ) -> Tensor:
        """Fused forward in compute dtype ``dt``.  ``residual`` is added after the last op (LayerNorm if present).
        ``first_gathers`` / ``first_cols`` feed the split first layer of GraphConv's edge MLP (gather-add epilogue)."""
        mods = list(self.mlp)
        i, first = 0, True
        while i < len(mods):
            lin = mods[i]
            act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight")
            last = i + (2 if act else 1) >= len(mods)
            kw = {}
            if first and first_gathers is not None:
                kw["gather1"], kw["gather2"] = first_gathers
            if last and self.layer_norm is None:
                kw["residual"], kw["out"] = residual, out
                kw["want_stats"] = want_stats  # the caller's next op is a LayerNorm folded into a GEMM: hand it the row statistics
            if first and pre_ln is not None:  # LayerNorm(x) feeding the first Linear: folded into that GEMM on the bf16 path
                x = Fn.ln_linear(self._pack, x, pre_ln, ("pre_ln", id(lin)), Fn.linear_sources([lin]), lambda lin=lin: Fn.cat_linear32([lin]), dt,
                                 gelu=act, **kw)
            else:
                x = Fn.fused_linear(self._pack, x, [lin], dt, cols=first_cols if first else None, gelu=act, **kw)
            i += 2 if act else 1
            first = False
        if self.layer_norm is not None:
            from .normalization import _check_plain_layernorm

            _check_plain_layernorm(self.layer_norm)
            from .. import ops

            x = ops.layer_norm(x, self._pack.f32(self.layer_norm.weight), self._pack.f32(self.layer_norm.bias), self.layer_norm.eps,
                               residual=residual, out=out, out_dtype=dt)  # fmt: skip
        return x

    def forward(self, x: Tensor, **layer_kwargs) -> Tensor:
        if layer_kwargs:
            raise NotImplementedError("conditional LayerNorm kwargs are not implemented")
        Fn.forward_only_guard(self)
        shape = x.shape
        y = self.run(x.reshape(-1, shape[-1]), Fn.compute_dtype(x))
        return y.reshape(*shape[:-1], y.shape[-1])
