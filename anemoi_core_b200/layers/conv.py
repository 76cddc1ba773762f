"""Graph message-passing operators (reference: layers/conv.py:29-81 GraphConv, :84-147 GraphTransformerConv).

No torch_geometric ``MessagePassing`` here: both operators run over the cached dst-sorted CSR on sm_100a kernels.
"""

from __future__ import annotations

from typing import Optional
from typing import Union

import torch
from torch import Tensor
from torch import nn

from .. import ops
from ..distributed.khop_edges import sort_edge_index_by_dst
from . import _functional as Fn
from .mlp import MLP
from .mlp import GatedMLPLayer

PairTensor = tuple[Tensor, Tensor]

import os  # noqa: E402

GC_PI_BF16 = os.environ.get("ANEMOI_B200_GC_PI_BF16", "1") != "0"  # the same for the dst-indexed term
GC_PJ_BF16 = os.environ.get("ANEMOI_B200_GC_PJ_BF16", "1") != "0"  # bf16 table for the src-indexed gather term of GraphConv (A/B switch)


def _sorted_plan(edge_index: Tensor, n_src: int, n_dst: int) -> tuple[ops.GraphCSR, Optional[Tensor]]:
    """CSR plan for ``edge_index``; if it turns out not to be dst-sorted, stable-sort it (like the reference's
    ``edge_index_to_csc(..., edges_are_dst_sorted=False)``, triton/utils.py:49-59) and return the int32 permutation."""
    try:
        return Fn.csr_for(edge_index, n_src, n_dst), None
    except ValueError as err:
        if "not sorted" not in str(err):
            raise
    sorted_ei, perm = sort_edge_index_by_dst(edge_index)
    return ops.build_csr(sorted_ei.contiguous(), n_src, n_dst), perm.to(torch.int32)


class GraphConv(Fn.PackOwner):
    """e' = edge_mlp([x_dst[dst], x_src[src], e]) + e ;  out[d] = sum_{edges into d} e'   (conv.py:66-81).

    Kernel decomposition: the first Linear of ``edge_mlp`` acts on a concatenation, so
    ``W1 [x_i; x_j; e] = (W1_i x_dst)[dst] + (W1_j x_src)[src] + W1_e e``: two node-level GEMMs (N rows instead of E)
    feed a gather-add epilogue of the edge-level GEMM, the [E, 3C] concatenation never exists; the LayerNorm, the
    ``+ e`` residual and the scatter-sum are one segmented-reduction kernel over the dst-sorted edges.
    """

    def __init__(self, in_channels: int, out_channels: int, layer_kernels=None, mlp_extra_layers: int = 0, mlp_implementation: str = "mlp", **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.edge_mlp = MLP(3 * in_channels, out_channels, out_channels, layer_kernels=layer_kernels, n_extra_layers=mlp_extra_layers + 1,
                            mlp_implementation=mlp_implementation)  # fmt: skip
        self._pack = Fn.WeightPack()

    def _fused_plan(self):
        """Linear layers of ``edge_mlp`` if the whole operator can run as the ONE-kernel form (``ops.graphconv_fused``): width 16 / 32 / 64,
        plain Linear -> GELU -> ... -> Linear (no gated layers), plain LayerNorm.  None otherwise.  ``ANEMOI_B200_GC_FUSED=0`` forces the
        decomposed form (A/B, tests)."""
        import os

        C, mlp = self.in_channels, self.edge_mlp
        if os.environ.get("ANEMOI_B200_GC_FUSED", "1") == "0" or C not in ops.GRAPHCONV_FUSED_WIDTHS:
            return None
        mods = list(mlp.mlp)
        lins = mods[0::2]
        if len(mods) % 2 == 0 or not 2 <= len(lins) <= ops.GRAPHCONV_FUSED_MAX_LAYERS:
            return None
        if not all(isinstance(m, nn.Linear) for m in lins) or not all(isinstance(m, nn.GELU) and m.approximate == "none" for m in mods[1::2]):
            return None
        if lins[0].weight.shape != (C, 3 * C) or any(m.weight.shape != (C, C) for m in lins[1:]):
            return None
        ln = mlp.layer_norm
        if not isinstance(ln, nn.LayerNorm) or tuple(ln.normalized_shape) != (C,):
            return None
        return lins

    def _run_fused(self, lins, x_src: Tensor, x_dst: Tensor, edge_attr: Tensor, csr: ops.GraphCSR, dt: torch.dtype, out: Optional[Tensor]):
        C, pack, ln = self.in_channels, self.edge_mlp._pack, self.edge_mlp.layer_norm
        ws, bs = [m.weight for m in lins], [m.bias for m in lins]
        w = pack.get(("gcf_w", dt), ws, lambda: torch.cat([x.detach().reshape(-1) for x in ws]).to(dt).contiguous())
        b = pack.get(("gcf_b",), bs, lambda: torch.stack([torch.zeros(C, device=ws[0].device) if x is None else x.detach().float() for x in bs]).contiguous())
        e = Fn.as_operand(edge_attr, dt, C)
        e_new, out = ops.graphconv_fused(Fn.as_operand(x_src, dt, C), Fn.as_operand(x_dst, dt, C), e, w, b, len(lins), pack.f32(ln.weight),
                                         pack.f32(ln.bias), csr, ln.eps, out=out)  # fmt: skip
        return out, e_new

    def run(self, x_src: Tensor, x_dst: Tensor, edge_attr: Tensor, csr: ops.GraphCSR, dt: torch.dtype, out: Optional[Tensor] = None):
        C = self.in_channels
        mlp = self.edge_mlp
        if mlp.layer_norm is None:
            raise NotImplementedError("GraphConv.edge_mlp without LayerNorm")
        lins = self._fused_plan()
        if lins is not None:
            return self._run_fused(lins, x_src, x_dst, edge_attr, csr, dt, out)
        def group(m):  # the Linear containers of one feed-forward layer: gate | value rows of a gated layer run as ONE GEMM
            return [m.gate_proj, m.value_proj] if isinstance(m, GatedMLPLayer) else [m]

        first = group(mlp.mlp[0])
        # node-level projections of the first edge-MLP layer, gathered and added in the edge GEMM's epilogue.  As fp32 tables they cost 8 C
        # bytes of L2 reads per edge and DOUBLED that GEMM (552 -> 1 099 us at C = 1024, profiles/r2/call51_*); on the bf16 path both tables
        # are bf16 (two extra roundings of two of the three addends: the GNN bf16 error moves from 4.66e-3 to 4.68e-3, bars unchanged):
        # 777 us, cfg3 step 110 -> 104.4 ms (call52_*, call53_*).  ANEMOI_B200_GC_PI_BF16=0 / _PJ_BF16=0 keep a table in fp32.
        pi_dt = dt if (dt == torch.bfloat16 and GC_PI_BF16) else torch.float32
        p_i = Fn.fused_linear(self._pack, x_dst, first, dt, cols=slice(0, C), use_bias=False, out_dtype=pi_dt)
        pj_dt = dt if (dt == torch.bfloat16 and GC_PJ_BF16) else torch.float32
        p_j = Fn.fused_linear(self._pack, x_src, first, dt, cols=slice(C, 2 * C), use_bias=False, out_dtype=pj_dt)
        e = Fn.as_operand(edge_attr, dt, edge_attr.shape[1])
        # hidden layers through the MLP runner up to (not including) the LayerNorm
        mods = list(mlp.mlp)
        h, i, is_first = e, 0, True
        while i < len(mods):
            lin = mods[i]
            gated = isinstance(lin, GatedMLPLayer)
            act = not gated and i + 1 < len(mods) and not hasattr(mods[i + 1], "weight")
            kw = {}
            if is_first:
                kw = {"gather1": (p_i, csr.dst32), "gather2": (p_j, csr.src32), "cols": slice(2 * C, 3 * C)}
            h = Fn.fused_linear(mlp._pack, h, group(lin), dt, gelu=act, **kw)
            if gated:
                h = ops.glu_combine(h, lin.kind)
            i += 2 if act else 1
            is_first = False
        ln = mlp.layer_norm
        e_new, out = ops.graphconv_ln_aggregate(h, mlp._pack.f32(ln.weight), mlp._pack.f32(ln.bias), e, csr, ln.eps, out=out)
        return out, e_new

    def forward(self, x: Union[Tensor, PairTensor], edge_attr: Tensor, edge_index: Tensor, size=None):
        x_src, x_dst = x if isinstance(x, (tuple, list)) else (x, x)
        dt = Fn.compute_dtype(x_src, x_dst, edge_attr)
        csr, perm = _sorted_plan(edge_index, x_src.shape[0], x_dst.shape[0])
        if perm is not None:
            raise ValueError("GraphConv returns per-edge features in edge order: pass a dst-sorted edge_index (sort_edge_index_by_dst)")
        from . import _train as T

        if T.wants_grad(self, x_src, x_dst, edge_attr):  # differentiable path (layers/_train.py)
            return T.graph_conv(self, x_src, x_dst, edge_attr, csr, dt)
        return self.run(x_src, x_dst, edge_attr, csr, dt)


class GraphTransformerConv(nn.Module):
    """out[d,h] = sum_e softmax_e(q[d,h].(k[src_e,h]+e[e,h]) / sqrt(Ch)) (v[src_e,h]+e[e,h])   (conv.py:103-147).

    Operator-level boundary of the reference (``self.conv(query, key, value, edges, edge_index, size)``,
    block.py:787-791) with the edge projection materialised; the blocks use the fused lin_edge form instead."""

    def __init__(self, out_channels: int, dropout: float = 0.0, **kwargs):
        super().__init__()
        if dropout:
            raise NotImplementedError("attention dropout is not implemented")
        self.out_channels = out_channels
        self.dropout = dropout

    def forward(self, query: Tensor, key: Tensor, value: Tensor, edge_attr: Optional[Tensor], edge_index: Tensor, size=None) -> Tensor:
        n_dst, heads, ch = query.shape
        n_src = key.shape[0]
        csr, perm = _sorted_plan(edge_index, n_src, n_dst)
        dt = query.dtype
        e = None
        if edge_attr is not None:
            e = edge_attr.reshape(edge_attr.shape[0], heads * ch)
            e = ops.cast_pad(e, dt, idx=perm) if (perm is not None or e.dtype != dt) else e
        q2, k2, v2 = query.reshape(n_dst, heads * ch), key.reshape(n_src, heads * ch), value.reshape(n_src, heads * ch)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (query, key, value, edge_attr)):
            from .. import autograd as AG

            if perm is not None:  # differentiable gather into the sorted order (the kernel's cast_pad above carries no gradient)
                e = edge_attr.reshape(edge_attr.shape[0], heads * ch).index_select(0, perm.long()).to(dt)
            elif e is not None and e.dtype != edge_attr.dtype:
                e = edge_attr.reshape(edge_attr.shape[0], heads * ch).to(dt)
            return AG.gt_attention(q2, k2, v2, e, csr, heads).reshape(n_dst, heads, ch)
        out = ops.gt_attention(q2, k2, v2, csr, heads, e_proj=e)
        return out.reshape(n_dst, heads, ch)
