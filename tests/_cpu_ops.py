"""TEST INFRASTRUCTURE ONLY — a plain-PyTorch fp32 stand-in for ``anemoi_core_b200.ops`` so that the multi-process HOST logic of the
sharded forward (edge sharding, halo plans and exchanges, gathers, relabelled CSR plans, shard bookkeeping across ranks) can be exercised
end to end under Gloo on a machine without a GPU.  It is never imported by the package; ``install()`` monkey-patches the ``ops`` module of the
current (test) process.  The kernels themselves are validated on the GPU (``-m gpu``) against the oracle; these tests only compare a sharded
run with a single-rank run of the SAME stand-in arithmetic, so the stand-in does not need to be (and is not) a second oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass
class _CSR:
    n_src: int
    n_dst: int
    n_edges: int
    colptr: Tensor
    colptr32: Tensor
    src32: Tensor
    dst32: Tensor
    rev: Optional[tuple] = None


def build_csr(edge_index: Tensor, n_src: int, n_dst: int, validate: bool = True) -> _CSR:
    src, dst = edge_index[0], edge_index[1]
    if validate and edge_index.shape[1]:
        if bool((dst[1:] < dst[:-1]).any()):
            raise ValueError("edge_index is not sorted by destination")
        if int(src.min()) < 0 or int(src.max()) >= n_src or int(dst.min()) < 0 or int(dst.max()) >= n_dst:
            raise ValueError("edge_index out of range")
    colptr = torch.zeros(n_dst + 1, dtype=torch.int64)
    colptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)
    return _CSR(n_src, n_dst, edge_index.shape[1], colptr, colptr.int(), src.int().contiguous(), dst.int().contiguous())


def _block_stats(y: Tensor) -> Tensor:
    M, N = y.shape
    P = (N + 63) // 64
    out = torch.zeros(M, P, 2)
    for b in range(P):
        blk = y[:, b * 64 : min(N, (b + 1) * 64)].float()
        mean = blk.mean(1)
        out[:, b, 0], out[:, b, 1] = mean, ((blk - mean[:, None]) ** 2).sum(1)
    return out


def partial_stats_buffer(out_rows: int, out_cols: int, device) -> Tensor:
    return torch.empty((out_rows, (out_cols + 63) // 64, 2), dtype=torch.float32)


def linear(a, weight, bias=None, gelu=False, residual=None, gather1=None, gather2=None, out=None, out_dtype=None, ln_stats=None, ln_colsum=None,
           ln_dim=0, ln_eps=0.0, stats_out=None):  # fmt: skip
    acc = a.float() @ weight.float().t()
    if ln_stats is not None:
        if ln_stats.dim() == 3:
            P = ln_stats.shape[1]  # per 64-column block (mean, M2): Chan merge, like common.cuh:ln_row_mean_rstd
            n = torch.full((P,), 64.0)
            n[-1] = ln_dim - 64 * (P - 1)
            mean = (ln_stats[..., 0] * n).sum(1) / ln_dim
            m2 = (ln_stats[..., 1] + n * (ln_stats[..., 0] - mean[:, None]) ** 2).sum(1)
            rstd = (m2 / ln_dim + ln_eps).rsqrt()
        else:
            mean, rstd = ln_stats[:, 0], ln_stats[:, 1]
        acc = rstd[:, None] * (acc - mean[:, None] * ln_colsum[None, :])
    if bias is not None:
        acc = acc + bias
    for g in (gather1, gather2):
        if g is not None:
            acc = acc + g[0][g[1].long(), : acc.shape[1]]
    if gelu:
        acc = F.gelu(acc)
    if residual is not None:
        acc = acc + residual.float()
    res = acc.to(out.dtype if out is not None else (out_dtype or a.dtype))
    if out is not None:
        out.copy_(res)
        res = out
    if stats_out is not None:
        stats_out.copy_(_block_stats(res))
    return res


def layer_norm(x, weight, bias, eps=1e-5, residual=None, out=None, out_dtype=None, groups=1):
    M, W = x.shape
    C = W // groups
    y = F.layer_norm(x.float().reshape(M, groups, C), (C,), weight, bias, eps).reshape(M, W)
    if residual is not None:
        y = y + residual.float()
    y = y.to(out.dtype if out is not None else (out_dtype or x.dtype))
    if out is not None:
        out.copy_(y)
        return out
    return y


def row_stats(x, eps=1e-5):
    xf = x.float()
    return torch.stack([xf.mean(1), (xf.var(1, unbiased=False) + eps).rsqrt()], 1)


def gt_attention(q, k, v, csr, heads, e_proj=None, edge_attr=None, w_edge=None, b_edge=None, qw=None, abar=None, dp=0, add=None, out=None, tiles=None,
                 lse=None):
    n_dst, C = q.shape
    Ch = C // heads
    src, dst = csr.src32.long(), csr.dst32.long()
    qh, kh, vh = q.float().view(n_dst, heads, Ch), k.float().view(-1, heads, Ch), v.float().view(-1, heads, Ch)
    ke, ve = kh[src], vh[src]
    folded = edge_attr is not None and qw is not None
    if e_proj is not None:
        e = e_proj.float().view(-1, heads, Ch)
        ke, ve = ke + e, ve + e
    elif edge_attr is not None and not folded:
        d = w_edge.shape[1]
        e = (edge_attr[:, :d] @ w_edge.t() + (b_edge if b_edge is not None else 0.0)).view(-1, heads, Ch)
        ke, ve = ke + e, ve + e
    score = (qh[dst] * ke).sum(-1) / (Ch**0.5)
    if folded:
        d = min(dp, edge_attr.shape[1])
        qwh = qw.float()[:, : heads * dp].reshape(n_dst, heads, dp)[:, :, :d]
        score = score + (qwh[dst] * edge_attr[:, None, :d]).sum(-1) / (Ch**0.5)
    mx = torch.full((n_dst, heads), -float("inf")).scatter_reduce(0, dst[:, None].expand(-1, heads), score, "amax", include_self=True)
    w = torch.exp(score - mx[dst])
    den = torch.zeros(n_dst, heads).index_add_(0, dst, w)
    alpha = w / den[dst]
    if lse is not None:  # natural-log normaliser per (dst, head); 0 for rows without edges (include/anemoi_b200.h)
        lse.copy_(torch.where(den > 0, mx + torch.log(den.clamp_min(1e-38)), torch.zeros_like(den)))
    res = torch.zeros(n_dst, heads, Ch).index_add_(0, dst, alpha[..., None] * ve)
    has = torch.zeros(n_dst, dtype=torch.bool)
    has[dst] = True
    if folded:
        if b_edge is not None:
            res = res + has[:, None, None] * b_edge.view(heads, Ch)
        ab = torch.zeros(n_dst, heads, dp)
        ab[:, :, :d] = torch.zeros(n_dst, heads, d).index_add_(0, dst, alpha[..., None] * edge_attr[:, None, :d])
        abar[:, : heads * dp] = ab.reshape(n_dst, heads * dp).to(abar.dtype)
    res = res.reshape(n_dst, C)
    if add is not None:
        res = res + add.float()
    res = res.to(q.dtype)
    if out is not None:
        out.copy_(res)
        return out
    return res


def graphconv_ln_aggregate(h, weight, bias, e, csr, eps=1e-5, out=None):
    e_new = (F.layer_norm(h.float(), (h.shape[1],), weight, bias, eps) + e.float()).to(h.dtype)
    agg = torch.zeros(csr.n_dst, h.shape[1]).index_add_(0, csr.dst32.long(), e_new.float()).to(h.dtype)
    if out is not None:
        out.copy_(agg)
        agg = out
    return e_new, agg


def graphconv_fused(x_src, x_dst, e, weights, biases, n_layers, gamma, beta, csr, eps=1e-5, out=None):
    C = e.shape[1]
    w = weights.float()
    h = torch.cat([x_dst.float()[csr.dst32.long()], x_src.float()[csr.src32.long()], e.float()], 1)
    off = 0
    for l in range(n_layers):
        k = 3 * C if l == 0 else C
        h = h @ w[off : off + C * k].reshape(C, k).t() + biases.reshape(n_layers, C)[l]
        off += C * k
        if l + 1 < n_layers:
            h = F.gelu(h)
    return graphconv_ln_aggregate(h.to(e.dtype), gamma, beta, e, csr, eps, out)


def cast_pad(x, dtype, k_pad=None, idx=None, out=None):
    src = x if idx is None else x[idx.long()]
    k_pad = src.shape[1] if k_pad is None else k_pad
    res = torch.zeros((src.shape[0], k_pad), dtype=dtype)
    res[:, : src.shape[1]] = src.to(dtype)
    if out is not None:
        out[:, :k_pad].copy_(res)
        return out
    return res


def add(a, b, out_dtype=None):
    return (a.float() + b.float()).to(out_dtype or a.dtype)


def assemble_input(x, attrs, out_dtype, k_pad=None):
    b, t, e, g, v = x.shape
    rows = x.permute(0, 2, 3, 1, 4).reshape(b * e * g, t * v)
    if attrs is not None:
        rows = torch.cat([rows, attrs.float().repeat(rows.shape[0] // attrs.shape[0], 1)], -1)
    out = torch.zeros((rows.shape[0], k_pad or rows.shape[1]), dtype=out_dtype)
    out[:, : rows.shape[1]] = rows.to(out_dtype)
    return out


def assemble_output(dec, x, batch, ensemble, n_step_output, step=-1, skip_src=None, bound=None):
    g = dec.shape[0] // (batch * ensemble)
    y = dec.float().reshape(batch, ensemble, g, n_step_output, -1).permute(0, 3, 1, 2, 4).clone()
    if x is not None and skip_src is not None:
        sel = skip_src >= 0
        y[..., sel] += x[:, step].unsqueeze(1)[..., skip_src[sel].long()]
    if bound is not None:
        y[..., bound == 1] = F.relu(y[..., bound == 1])
        y[..., bound == 2] = F.leaky_relu(y[..., bound == 2])
    return y


def glu_combine(gv, act):
    H = gv.shape[1] // 2
    gate = {"glu": torch.sigmoid, "swiglu": F.silu, "geglu": F.gelu, "reglu": torch.relu}[act]
    return (gate(gv[:, :H].float()) * gv[:, H:].float()).to(gv.dtype)


def cond_layer_norm(x, cond, w_scale, b_scale, w_bias, b_bias, eps=1e-5, out_dtype=None):
    y = F.layer_norm(x.float(), (x.shape[1],), None, None, eps)
    return (y * (1.0 + cond.float() @ w_scale.t() + b_scale) + cond.float() @ w_bias.t() + b_bias).to(out_dtype or x.dtype)


# ---- backward entry points (training on a model-parallel group under Gloo: tests/test_sharded_training_gloo.py) ---------------------
# Each one differentiates the stand-in forward above with PyTorch autograd: what is under test is the distributed autograd plumbing
# (HaloExchangeFn / GatherRowsFn / AllToAllFn, per-rank partial parameter gradients), not these formulas.
def gelu(x, dy=None):
    if dy is None:
        return F.gelu(x.float()).to(x.dtype)
    xf = x.detach().float().requires_grad_()
    with torch.enable_grad():
        y = F.gelu(xf)
    return torch.autograd.grad(y, xf, dy.float())[0].to(x.dtype)


def layer_norm_bwd(x, gamma, dy, eps, groups=1, dz=None, idx=None, want_dres=False):
    M, W = x.shape
    C = W // groups
    g = torch.zeros(M, W)
    if dy is not None:
        g = g + dy.float()
    if dz is not None:
        g = g + (dz.float()[idx.long()] if idx is not None else dz.float())
    xf = x.detach().float().requires_grad_()
    gm = (gamma.detach().float() if gamma is not None else torch.ones(C)).requires_grad_()
    bt = torch.zeros(C, requires_grad=True)
    with torch.enable_grad():
        y = F.layer_norm(xf.reshape(M, groups, C), (C,), gm, bt, eps).reshape(M, W)
    dx, dg, db = torch.autograd.grad(y, (xf, gm, bt), g)
    return dx.to(x.dtype), dg, db, (g.to(x.dtype) if want_dres else None)


def gt_attention_bwd(q, k, v, e_proj, out, dout, lse, csr, heads):
    ts = [t.detach().float().requires_grad_() for t in (q, k, v)]
    ef = None if e_proj is None else e_proj.detach().float().requires_grad_()
    with torch.enable_grad():
        y = gt_attention(ts[0], ts[1], ts[2], csr, heads, e_proj=ef)
    gs = torch.autograd.grad(y, ts + ([ef] if ef is not None else []), dout.float())
    de = gs[3].to(q.dtype) if ef is not None else None
    return gs[0].to(q.dtype), gs[1].to(q.dtype), gs[2].to(q.dtype), de


def glu_combine_bwd(gv, dy, act):
    gf = gv.detach().float().requires_grad_()
    with torch.enable_grad():
        y = glu_combine(gf, act)
    return torch.autograd.grad(y, gf, dy.float())[0].to(gv.dtype)


def segment_sum(rows, ptr32, eid32, n_out):
    ptr = ptr32.long()
    seg = torch.repeat_interleave(torch.arange(n_out), ptr[1:] - ptr[:-1])
    sel = rows.float()[eid32.long()] if eid32 is not None else rows.float()[: int(ptr[-1])]
    return torch.zeros(n_out, rows.shape[1]).index_add_(0, seg, sel).to(rows.dtype)


def col_sum(x):
    return x.float().sum(0)


def install() -> None:
    """Replace the CUDA entry points of ``anemoi_core_b200.ops`` in THIS process (a spawned Gloo test worker)."""
    import anemoi_core_b200.layers._functional as Fn
    from anemoi_core_b200 import ops

    for name in ("build_csr", "linear", "layer_norm", "row_stats", "gt_attention", "graphconv_ln_aggregate", "graphconv_fused", "cast_pad", "add", "partial_stats_buffer",
                 "assemble_input", "assemble_output", "glu_combine", "cond_layer_norm", "gelu", "layer_norm_bwd", "gt_attention_bwd", "glu_combine_bwd",
                 "segment_sum", "col_sum"):
        setattr(ops, name, globals()[name])
    ops.attention_tiles = lambda csr: None  # the tile plan only feeds the CUDA kernel
    ops._need_cuda = lambda *a, **k: None
    torch.cuda.is_current_stream_capturing = lambda: False  # csr_for asks; there is no CUDA runtime here
    assert Fn.ops is ops
