"""MLP container + fused forward (reference: layers/mlp.py:97-179).

Same structure and parameter names as the reference — ``mlp`` = Sequential(Linear, act, [Linear, act]*k, Linear[, act]),
or with a gated ``mlp_implementation`` (glu / swiglu / geglu / reglu) Sequential(GatedMLPLayer, [GatedMLPLayer]*k, Linear) —
optional ``layer_norm`` — so reference ``state_dict``s load unchanged.  The forward issues one fused
GEMM(+bias+GELU) kernel per Linear and one LayerNorm(+residual) kernel.
"""

from __future__ import annotations

import os

from typing import Optional

import torch
from torch import Tensor
from torch import nn

from . import _functional as Fn
from .utils import load_layer_kernels


def _is_gelu(m: nn.Module) -> bool:
    return isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none"


GATED = ("glu", "swiglu", "geglu", "reglu")
# Row-block size of MLP.run in MB of hidden tensor (0 = never chunk, the default); see MLP.run.  Measured (profiles/r2/call22_*): keeping the
# 168 MB hidden tensor of a cfg2 MLP in L2 by running the two GEMMs on wave-aligned blocks of 9 472 rows makes the step SLOWER - 9.47 ms
# against 8.59 ms (8.83 ms with two-wave blocks) - because every launch pays its own pipeline fill and an un-overlapped last epilogue
# (~10 us on a 20 us GEMM).  The L2 win needs both GEMMs in ONE persistent kernel, which TMEM cannot hold.
MLP_CHUNK_MB = int(os.environ.get("ANEMOI_B200_MLP_CHUNK_MB", "0"))


class GatedMLPLayer(nn.Module):
    """``gating(gate_proj(x)) * value_proj(x)`` (layers/mlp.py:38-53), same parameter names.  Forward: ONE GEMM on the row-concatenated
    gate | value weights (the folded LayerNorm and the row statistics work as for a plain Linear), then ``ops.glu_combine``."""

    def __init__(self, in_features: int, out_features: int, layer_kernels, mlp_implementation: str) -> None:
        super().__init__()
        self.gate_proj = layer_kernels.Linear(in_features, out_features)
        self.value_proj = layer_kernels.Linear(in_features, out_features)
        self.gating = {"glu": nn.Sigmoid, "swiglu": nn.SiLU, "geglu": nn.GELU, "reglu": nn.ReLU}[mlp_implementation]()
        self.kind = mlp_implementation


class MLP(Fn.PackOwner):
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

    def _out_features(self) -> int:
        last = [m for m in self.mlp if hasattr(m, "weight")][-1]
        return last.weight.shape[0]

    def _chunk_rows(self, x: Tensor, dt: torch.dtype, first_gathers, want_stats: bool) -> int:
        """Rows per block for ``run`` (0 = no chunking): only for bf16 chains without gather-add / statistics hand-over whose widest hidden
        tensor exceeds MLP_CHUNK_MB.  A block is a whole number of GEMM WAVES of the last (narrowest) GEMM - 256-row tiles, pairs of SMs, e.g.
        37 row tiles x 2 column tiles = 74 CTA pairs for N = 512 - so that chunking adds no tile-quantisation loss: for the cfg2 MLP
        (512 -> 2048 -> 512) a block of 9 472 rows is exactly 4 waves of the first GEMM and 1 wave of the second, 38.8 MB of hidden tensor."""
        if MLP_CHUNK_MB <= 0 or dt != torch.bfloat16 or first_gathers is not None or (want_stats and Fn.FUSED_ROW_STATS) or x.dim() != 2 or not x.is_cuda:
            return 0
        widths = [(2 * m.gate_proj.weight.shape[0]) if isinstance(m, GatedMLPLayer) else m.weight.shape[0] for m in self.mlp
                  if hasattr(m, "weight") or isinstance(m, GatedMLPLayer)]  # fmt: skip
        if len(widths) < 2:
            return 0
        M, hidden = x.shape[0], max(widths[:-1])
        if M * hidden * 2 <= (MLP_CHUNK_MB << 20) * 3 // 2:
            return 0
        pairs = max(1, torch.cuda.get_device_properties(x.device).multi_processor_count // 2)
        wave_tiles = max(1, pairs // (-(-widths[-1] // 256)))  # row tiles of the last GEMM that fill one wave of CTA pairs
        wave_bytes = wave_tiles * 256 * hidden * 2
        rows = wave_tiles * 256 * max(1, (MLP_CHUNK_MB << 20) // wave_bytes)
        return rows if rows < M else 0

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
        _chunked: bool = False,
    ) -> Tensor:
        """Fused forward in compute dtype ``dt``.  ``residual`` is added after the last op (LayerNorm if present).
        ``first_gathers`` / ``first_cols`` feed the split first layer of GraphConv's edge MLP (gather-add epilogue).
        ``serpentine``: L2-aware traversal (``ops.set_traversal``) for an input the previous GEMM has just written top-down: the row
        statistics of ``pre_ln`` walk bottom-up, the first GEMM top-down, the last GEMM (which reads the hidden tensor the first one has
        just written) bottom-up."""
        from .. import ops as _ops

        # Row chunking: the hidden tensor of a wide MLP ([40 962, 2048] bf16 = 168 MB in a cfg2 layer) is written by one GEMM and read back by
        # the next; cut into row blocks of <= MLP_CHUNK_MB the pair of GEMMs works on a hidden block that is still in L2 (126 MB) when it is
        # consumed, instead of streaming it through HBM twice.  Every op of the chain is row-wise, so the blocks are independent.
        rows = 0 if _chunked else self._chunk_rows(x, dt, first_gathers, want_stats)
        if rows:
            M = x.shape[0]
            if out is None:
                out = torch.empty((M, self._out_features()), dtype=dt, device=x.device)
            stats = None
            if pre_ln is not None and Fn.can_fold_ln(pre_ln, x.shape[1], dt) and x.dtype == dt and x.stride(1) == 1:
                if serpentine:
                    _ops.set_traversal(stats=True)
                stats = _ops.row_stats(x, pre_ln.eps)  # ONE statistics pass for all blocks; each block's view is tagged with its rows
                _ops.set_traversal()
            for r0 in range(0, M, rows):
                r1 = min(M, r0 + rows)
                xc = x[r0:r1]
                if stats is not None:
                    Fn.tag_row_stats(xc, stats[r0:r1])
                self.run(xc, dt, residual=None if residual is None else residual[r0:r1], first_cols=first_cols, out=out[r0:r1], pre_ln=pre_ln,
                         cond=None if cond is None else cond[r0:r1], serpentine=serpentine, _chunked=True)  # fmt: skip
            return out
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
