"""MLP container + fused forward (reference: layers/mlp.py:97-179).

Same structure and parameter names as the reference — ``mlp`` = Sequential(Linear, act, [Linear, act]*k, Linear[, act]),
or with a gated ``mlp_implementation`` (glu / swiglu / geglu / reglu) Sequential(GatedMLPLayer, [GatedMLPLayer]*k, Linear) —
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


GATED = ("glu", "swiglu", "geglu", "reglu")


class GatedMLPLayer(nn.Module):
    """``gating(gate_proj(x)) * value_proj(x)`` (layers/mlp.py:38-53), same parameter names.  Forward: ONE GEMM on the row-concatenated
    gate | value weights (the folded LayerNorm and the row statistics work as for a plain Linear), then ``ops.glu_combine``."""

    def __init__(self, in_features: int, out_features: int, layer_kernels, mlp_implementation: str) -> None:
        super().__init__()
        self.gate_proj = layer_kernels.Linear(in_features, out_features)
        self.value_proj = layer_kernels.Linear(in_features, out_features)
        self.gating = {"glu": nn.Sigmoid, "swiglu": nn.SiLU, "geglu": nn.GELU, "reglu": nn.ReLU}[mlp_implementation]()
        self.kind = mlp_implementation


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
        if mlp_implementation not in ("mlp",) + GATED:
            raise ValueError(f"`mlp_implementation` must be one of {('mlp',) + GATED}, got '{mlp_implementation}'.")
        self.mlp_implementation = mlp_implementation
        k = load_layer_kernels(layer_kernels)
        gated = mlp_implementation != "mlp"

        def ffn(i: int, o: int) -> list[nn.Module]:  # layers/mlp.py:70-94 build_feedforward_modules
            return [GatedMLPLayer(i, o, k, mlp_implementation)] if gated else [k.Linear(i, o), k.Activation()]

        layers: list[nn.Module] = ffn(in_features, hidden_dim)
        for _ in range(n_extra_layers):
            layers += ffn(hidden_dim, hidden_dim)
        layers.append(k.Linear(hidden_dim, out_features))
        if final_activation:
            if gated:
                raise NotImplementedError("final_activation with a gated mlp_implementation (no in-scope caller sets it, mapper.py:1052)")
            layers.append(k.Activation())
        for m in layers:
            if not hasattr(m, "weight") and not isinstance(m, GatedMLPLayer) and not _is_gelu(m):
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
        cond: Optional[Tensor] = None,
        serpentine: bool = False,
    ) -> Tensor:
        """Fused forward in compute dtype ``dt``.  ``residual`` is added after the last op (LayerNorm if present).
        ``first_gathers`` / ``first_cols`` feed the split first layer of GraphConv's edge MLP (gather-add epilogue).
        ``serpentine``: L2-aware traversal (``ops.set_traversal``) for an input the previous GEMM has just written top-down: the row
        statistics of ``pre_ln`` walk bottom-up, the first GEMM top-down, the last GEMM (which reads the hidden tensor the first one has
        just written) bottom-up."""
        from .. import ops as _ops

        mods = list(self.mlp)
        i, first = 0, True
        while i < len(mods):
            lin = mods[i]
            if isinstance(lin, GatedMLPLayer):  # never the last module: the final Linear follows
                from .. import ops

                pair = [lin.gate_proj, lin.value_proj]
                kw = {}
                if first and first_gathers is not None:
                    kw["gather1"], kw["gather2"] = first_gathers
                if first and pre_ln is not None:
                    gv = Fn.ln_linear(self._pack, x, pre_ln, ("pre_ln_gated", id(lin)), Fn.linear_sources(pair), lambda pair=pair: Fn.cat_linear32(pair),
                                      dt, cond=cond, **kw)  # fmt: skip
                else:
                    gv = Fn.fused_linear(self._pack, x, pair, dt, cols=first_cols if first else None, **kw)
                x = ops.glu_combine(gv, lin.kind)
                i += 1
                first = False
                continue
            act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight")
            last = i + (2 if act else 1) >= len(mods)
            kw = {}
            if first and first_gathers is not None:
                kw["gather1"], kw["gather2"] = first_gathers
            if last and self.layer_norm is None:
                kw["residual"], kw["out"] = residual, out
                kw["want_stats"] = want_stats  # the caller's next op is a LayerNorm folded into a GEMM: hand it the row statistics
            if serpentine:
                _ops.set_traversal(gemm=last and not first, stats=first)
            if first and pre_ln is not None:  # LayerNorm(x) feeding the first Linear: folded into that GEMM on the bf16 path
                x = Fn.ln_linear(self._pack, x, pre_ln, ("pre_ln", id(lin)), Fn.linear_sources([lin]), lambda lin=lin: Fn.cat_linear32([lin]), dt,
                                 cond=cond, gelu=act, **kw)
            else:
                x = Fn.fused_linear(self._pack, x, [lin], dt, cols=first_cols if first else None, gelu=act, **kw)
            i += 2 if act else 1
            first = False
        if serpentine:
            _ops.set_traversal()
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
        shape = x.shape
        from . import _train as T

        if T.wants_grad(self, x):  # differentiable path (layers/_train.py)
            y = T.mlp(self, x.reshape(-1, shape[-1]), Fn.compute_dtype(x))
            return y.reshape(*shape[:-1], y.shape[-1])
        y = self.run(x.reshape(-1, shape[-1]), Fn.compute_dtype(x))
        return y.reshape(*shape[:-1], y.shape[-1])
