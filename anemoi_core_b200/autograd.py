"""``torch.autograd.Function`` wiring of the C-ABI ops (backward of the fused ops, so the drop-in modules train).

Reference: the autograd halves the reference gets for free from PyTorch / PyG, and for its fused attention op
``triton/gt.py:451-556`` (``GraphTransformerFunction`` / ``register_autograd``) with the backward kernels ``gt.py:182-376``.

Forward of every Function is the same sm_100a kernel the inference path runs; backward:
  LinearFn          dx = dz W on the tcgen05 GEMM (``ops.linear`` with the transposed weight), dW = dz^T x as a plain library GEMM
                    (cuBLAS through ``torch.matmul``: a bare GEMM with nothing to fuse), db = column sum; GELU through ``ops.gelu``.
  LayerNormFn       ``ops.layer_norm_bwd``.
  GTAttentionFn     ``ops.gt_attention_bwd`` (dst-major + src-major passes over the cached CSR / reverse CSR; deterministic).
  GraphConvTailFn   ``ops.layer_norm_bwd`` with the gathered second cotangent (d out[dst[i]]): e' = LN(h) + e, out = segment-sum(e').
No CPU path: CPU tensors raise in ``ops`` like everywhere else.
"""

from __future__ import annotations

import os
from typing import Optional

import torch
from torch import Tensor
from torch.autograd import Function

from . import ops


COL_SUM = os.environ.get("ANEMOI_B200_COL_SUM", "1") != "0"  # bias gradients through ops.col_sum instead of PyTorch's sum(0)


def _bias_grad(dz: Tensor) -> Tensor:
    """db = column sums of the cotangent, fp32 accumulation without an fp32 copy of dz."""
    if COL_SUM and dz.dim() == 2 and dz.stride(1) == 1 and dz.data_ptr() % 16 == 0 and (dz.stride(0) * dz.element_size()) % 16 == 0:
        return ops.col_sum(dz)
    return dz.sum(0, dtype=torch.float32)


def _pad_cols(t: Tensor, k: int) -> Tensor:
    return t if t.shape[1] == k else torch.nn.functional.pad(t, (0, k - t.shape[1]))


class LinearFn(Function):
    """y = [gelu](x W^T + b) [+ residual] in compute dtype ``dt`` (x [M, K] any float dtype, W [N, K], b [N] | None, residual [M, N] | None: added
    in the GEMM epilogue like in inference; its gradient is the cotangent itself)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], gelu: bool, dt: torch.dtype, residual: Optional[Tensor] = None) -> Tensor:
        K = x.shape[1]
        kp = max(64, (K + 7) // 8 * 8) if dt == torch.bfloat16 else K  # tcgen05 operands: 16-byte rows, one swizzle span
        xd = x.detach()
        # no copy when x already is a dt operand of the right width (the common case inside a block: the previous op's output)
        xa = xd if (xd.dtype == dt and K == kp and xd.stride(1) == 1 and xd.data_ptr() % 16 == 0 and (xd.stride(0) * xd.element_size()) % 16 == 0) else ops.cast_pad(xd, dt, kp)
        wa = _pad_cols(weight.detach(), kp).to(dt).contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        ctx.res_dt = None if residual is None else residual.dtype
        if residual is not None and not gelu:
            r = residual.detach()
            z = ops.linear(xa, wa, b32, residual=r if (r.dtype == dt and r.stride(1) == 1) else r.to(dt).contiguous())
        else:
            z = ops.linear(xa, wa, b32)
        ctx.save_for_backward(xa, wa, z if gelu else None)
        ctx.meta = (K, x.dtype, weight.dtype, None if bias is None else bias.dtype, gelu)
        y = ops.gelu(z) if gelu else z
        return y + residual.detach().to(dt) if (residual is not None and gelu) else y

    @staticmethod
    def backward(ctx, dy: Tensor):
        xa, wa, z = ctx.saved_tensors
        K, x_dt, w_dt, b_dt, gelu = ctx.meta
        dy = dy.to(xa.dtype)
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        dz = ops.gelu(z, dy) if gelu else dy
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear(dz, wa.t().contiguous())[:, :K].to(x_dt)
        if ctx.needs_input_grad[1]:
            dw = torch.matmul(dz.t(), xa)[:, :K].to(w_dt)  # plain library GEMM (weight gradient; reduction over all rows)
        if b_dt is not None and ctx.needs_input_grad[2]:
            db = _bias_grad(dz).to(b_dt)
        dres = dy.to(ctx.res_dt) if (ctx.res_dt is not None and ctx.needs_input_grad[5]) else None
        return dx, dw, db, None, None, dres


class EdgeFirstLayerFn(Function):
    """z = e W_e^T + b + p_i[dst] + p_j[src]: the first layer of GraphConv's edge MLP on the concatenation [x_i; x_j; e] with the two node-level
    projections gathered in the GEMM epilogue (forward: no [E, C] gather is materialised).  Backward: dz -> the GEMM (de, dW_e, db) and,
    without atomics, dp_i = segment sums of dz over the dst-sorted edge runs, dp_j = segment sums over the reverse CSR (``ops.segment_sum``;
    PyTorch's index_add_ took 131 of the 520 ms of a cfg3 training step)."""

    @staticmethod
    def forward(ctx, e: Tensor, w_e: Tensor, bias: Optional[Tensor], p_i: Tensor, p_j: Tensor, csr: ops.GraphCSR, dt: torch.dtype) -> Tensor:
        K = e.shape[1]
        kp = max(64, (K + 7) // 8 * 8) if dt == torch.bfloat16 else K
        ed = e.detach()
        ea = ed if (ed.dtype == dt and K == kp and ed.stride(1) == 1 and ed.data_ptr() % 16 == 0 and (ed.stride(0) * ed.element_size()) % 16 == 0) else ops.cast_pad(ed, dt, kp)
        wa = _pad_cols(w_e.detach(), kp).to(dt).contiguous()
        z = ops.linear(ea, wa, None if bias is None else bias.detach().float().contiguous(),
                       gather1=(p_i.detach().float().contiguous(), csr.dst32), gather2=(p_j.detach().float().contiguous(), csr.src32))
        ctx.save_for_backward(ea, wa)
        ctx.csr = csr
        ctx.meta = (K, e.dtype, w_e.dtype, None if bias is None else bias.dtype, p_i.dtype, p_j.dtype)
        return z

    @staticmethod
    def backward(ctx, dz: Tensor):
        ea, wa = ctx.saved_tensors
        K, e_dt, w_dt, b_dt, pi_dt, pj_dt = ctx.meta
        csr = ctx.csr
        dz = dz.to(ea.dtype)
        dz = dz if (dz.stride(1) == 1 and dz.data_ptr() % 16 == 0) else dz.contiguous()
        de = dw = db = dpi = dpj = None
        if ctx.needs_input_grad[0]:
            de = ops.linear(dz, wa.t().contiguous())[:, :K].to(e_dt)
        if ctx.needs_input_grad[1]:
            dw = torch.matmul(dz.t(), ea)[:, :K].to(w_dt)
        if b_dt is not None and ctx.needs_input_grad[2]:
            db = _bias_grad(dz).to(b_dt)
        if ctx.needs_input_grad[3]:
            dpi = ops.segment_sum(dz, csr.colptr32, None, csr.n_dst).to(pi_dt)
        if ctx.needs_input_grad[4]:
            rev_ptr, rev_eid = ops.reverse_csr(csr)
            dpj = ops.segment_sum(dz, rev_ptr, rev_eid, csr.n_src).to(pj_dt)
        return de, dw, db, dpi, dpj, None, None


class GeluFn(Function):
    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        x = x.detach()
        x = x if x.stride(1) == 1 else x.contiguous()
        ctx.save_for_backward(x)
        return ops.gelu(x)

    @staticmethod
    def backward(ctx, dy: Tensor):
        (x,) = ctx.saved_tensors
        dy = dy.to(x.dtype)
        return ops.gelu(x, dy if dy.stride(1) == 1 else dy.contiguous())


class GluCombineFn(Function):
    """``act(gv[:, :H]) * gv[:, H:]`` of the gated feed-forward layers (layers/mlp.py:38-53)."""

    @staticmethod
    def forward(ctx, gv: Tensor, act: str) -> Tensor:
        gv = gv.detach()
        gv = gv if gv.stride(1) == 1 else gv.contiguous()
        ctx.save_for_backward(gv)
        ctx.act = act
        return ops.glu_combine(gv, act)

    @staticmethod
    def backward(ctx, dy: Tensor):
        (gv,) = ctx.saved_tensors
        dy = dy.to(gv.dtype)
        return ops.glu_combine_bwd(gv, dy if dy.stride(1) == 1 else dy.contiguous(), ctx.act), None


class LayerNormFn(Function):
    """LayerNorm over ``groups`` groups of the last dimension (weight / bias [C] | None), output in ``dt``."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float, groups: int, dt: torch.dtype) -> Tensor:
        x = x.detach()
        x = x if x.stride(1) == 1 else x.contiguous()
        w32 = None if weight is None else weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        ctx.save_for_backward(x, w32)
        ctx.meta = (eps, groups, None if weight is None else weight.dtype, None if bias is None else bias.dtype)
        return ops.layer_norm(x, w32, b32, eps, out_dtype=dt, groups=groups)

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w32 = ctx.saved_tensors
        eps, groups, w_dt, b_dt = ctx.meta
        dy = dy.to(x.dtype)
        dx, dg, db, _ = ops.layer_norm_bwd(x, w32, dy if dy.stride(1) == 1 else dy.contiguous(), eps, groups)
        return dx, (dg.to(w_dt) if w_dt is not None else None), (db.to(b_dt) if b_dt is not None else None), None, None, None


class GTAttentionFn(Function):
    """Edge-softmax attention with the materialised edge projection: the operator boundary of the reference
    (``anemoi::graph_transformer_attention``, triton/gt.py:390-447)."""

    @staticmethod
    def forward(ctx, q: Tensor, k: Tensor, v: Tensor, e: Optional[Tensor], csr: ops.GraphCSR, heads: int) -> Tensor:
        q, k, v = q.detach(), k.detach(), v.detach()
        e = None if e is None else e.detach()
        lse = torch.empty((q.shape[0], heads), dtype=torch.float32, device=q.device)
        out = ops.gt_attention(q, k, v, csr, heads, e_proj=e, lse=lse)
        ctx.save_for_backward(q, k, v, e, out, lse)
        ctx.csr, ctx.heads = csr, heads
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        q, k, v, e, out, lse = ctx.saved_tensors
        dout = dout.to(q.dtype)
        dq, dk, dv, de = ops.gt_attention_bwd(q, k, v, e, out, dout if dout.stride(1) == 1 else dout.contiguous(), lse, ctx.csr, ctx.heads)
        return dq, dk, dv, de, None, None


class GraphConvTailFn(Function):
    """(e', out) = (LayerNorm(h) * gamma + beta + e, segment-sum of e' over the dst-sorted edges): layers/conv.py:73-81."""

    @staticmethod
    def forward(ctx, h: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], e: Tensor, csr: ops.GraphCSR, eps: float):
        h, e = h.detach(), e.detach()
        h = h if h.stride(1) == 1 else h.contiguous()
        w32 = None if weight is None else weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        e_new, out = ops.graphconv_ln_aggregate(h, w32, b32, e.to(h.dtype), csr, eps)
        ctx.save_for_backward(h, w32)
        ctx.csr, ctx.meta = csr, (eps, None if weight is None else weight.dtype, None if bias is None else bias.dtype, e.dtype)
        ctx.set_materialize_grads(False)
        return e_new, out

    @staticmethod
    def backward(ctx, d_enew: Optional[Tensor], d_out: Optional[Tensor]):
        h, w32 = ctx.saved_tensors
        eps, w_dt, b_dt, e_dt = ctx.meta
        if d_enew is None and d_out is None:
            return None, None, None, None, None, None
        dy = None if d_enew is None else d_enew.to(h.dtype).contiguous()
        dz = None if d_out is None else d_out.to(h.dtype).contiguous()
        dh, dg, db, g = ops.layer_norm_bwd(h, w32, dy, eps, 1, dz=dz, idx=ctx.csr.dst32 if dz is not None else None, want_dres=True)
        return dh, (dg.to(w_dt) if w_dt is not None else None), (db.to(b_dt) if b_dt is not None else None), g.to(e_dt), None, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], dt: torch.dtype, gelu: bool = False, residual: Optional[Tensor] = None) -> Tensor:
    return LinearFn.apply(x, weight, bias, gelu, dt, residual)


def layer_norm(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float, dt: torch.dtype, groups: int = 1) -> Tensor:
    return LayerNormFn.apply(x, weight, bias, eps, groups, dt)


def edge_first_layer(e: Tensor, w_e: Tensor, bias: Optional[Tensor], p_i: Tensor, p_j: Tensor, csr: ops.GraphCSR, dt: torch.dtype) -> Tensor:
    return EdgeFirstLayerFn.apply(e, w_e, bias, p_i, p_j, csr, dt)


def glu_combine(gv: Tensor, act: str) -> Tensor:
    return GluCombineFn.apply(gv, act)


def gt_attention(q: Tensor, k: Tensor, v: Tensor, e: Optional[Tensor], csr: ops.GraphCSR, heads: int) -> Tensor:
    return GTAttentionFn.apply(q, k, v, e, csr, heads)


def graphconv_tail(h: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], e: Tensor, csr: ops.GraphCSR, eps: float):
    return GraphConvTailFn.apply(h, weight, bias, e, csr, eps)
