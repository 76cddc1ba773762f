"""Differentiable forward of the blocks / processors / mappers: the same sm_100a kernels wired through
``anemoi_core_b200.autograd``, composed like the reference modules (layers/block.py:362-395, 441-479, 963-1029, 1219-1273,
layers/mlp.py:158-179) so that PyTorch autograd reaches every parameter, the inputs and the (trainable) edge attributes.

Taken whenever gradients are needed (``wants_grad``); the inference path under ``torch.no_grad()`` keeps its fusions (LayerNorm folded
into the GEMMs, lin_edge folded into attention, packed weights), which have no backward.  On a model-parallel group the "edges" strategy
trains through the autograd halves of the exchange (``distributed.graph.HaloExchangeFn`` / ``GatherRowsFn``; reference
distributed/graph.py:227-500): the forward collective is NCCL, the backward its transpose; parameter gradients are per-rank partial sums
(the trainer's gradient all-reduce over the model group completes them, as in the reference).  Not implemented (NotImplementedError with
the reason): the heads strategy, gated MLP variants and ConditionalLayerNorm in training.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch import nn

from .. import autograd as AG
from . import _functional as Fn


def wants_grad(module: nn.Module, *tensors: Optional[Tensor]) -> bool:
    """Differentiable path iff autograd is recording and something upstream needs a gradient (a parameter in training mode or an input)."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return module.training and any(p.requires_grad for p in module.parameters())


def _single_gpu(group, what: str = "this module") -> None:
    from ..distributed.graph import group_size

    if group_size(group) > 1:
        raise NotImplementedError(f"training {what} on a model-parallel group is not implemented (the 'edges' strategy of the processors / mappers is)")


def _plain_ln(ln: nn.Module) -> None:
    from .normalization import _check_plain_layernorm

    _check_plain_layernorm(ln)


def norm(ln: nn.Module, x: Tensor, dt: torch.dtype, groups: int = 1) -> Tensor:
    _plain_ln(ln)
    return AG.layer_norm(x, ln.weight, getattr(ln, "bias", None), ln.eps, dt, groups)


def lin(layer: nn.Module, x: Tensor, dt: torch.dtype, gelu: bool = False) -> Tensor:
    return AG.linear(x, layer.weight, getattr(layer, "bias", None), dt, gelu)


def lin_cat(layers, x: Tensor, dt: torch.dtype) -> Tensor:
    """One GEMM for several Linear containers on the same input (torch.cat is differentiable: the gradients split back by themselves)."""
    w = torch.cat([l.weight for l in layers], 0)
    bs = [getattr(l, "bias", None) for l in layers]
    b = None if all(x_ is None for x_ in bs) else torch.cat([b_ if b_ is not None else torch.zeros(l.weight.shape[0], device=w.device) for b_, l in zip(bs, layers)])
    return AG.linear(x, w, b, dt)


def mlp(m, x: Tensor, dt: torch.dtype, residual: Optional[Tensor] = None, pre_ln: Optional[nn.Module] = None) -> Tensor:
    """``MLP.run`` with autograd: Linear(+GELU) chain, optional trailing LayerNorm, optional residual."""
    from .mlp import GatedMLPLayer

    if pre_ln is not None:
        x = norm(pre_ln, x, dt)
    mods = list(m.mlp)
    i = 0
    while i < len(mods):
        layer = mods[i]
        if isinstance(layer, GatedMLPLayer):
            raise NotImplementedError("training with gated MLP variants (glu / swiglu / geglu / reglu): backward of glu_combine is not implemented")
        act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight")
        x = lin(layer, x, dt, gelu=act)
        i += 2 if act else 1
    if m.layer_norm is not None:
        x = norm(m.layer_norm, x, dt)
    return x if residual is None else x + residual.to(x.dtype)


# ------------------------------------------------------------------------------------------------------------
# GraphConv (GNN)
# ------------------------------------------------------------------------------------------------------------
def graph_conv(conv, x_src: Tensor, x_dst: Tensor, e: Tensor, csr, dt: torch.dtype):
    """(out, e') of GraphConv (conv.py:66-81) with the split first layer: W1 [x_i; x_j; e] = (W1_i x_dst)[dst] + (W1_j x_src)[src] + W1_e e."""
    from .mlp import GatedMLPLayer

    m = conv.edge_mlp
    mods = list(m.mlp)
    if isinstance(mods[0], GatedMLPLayer) or m.layer_norm is None:
        raise NotImplementedError("training GraphConv with a gated edge MLP / without its LayerNorm")
    C = conv.in_channels
    w1, b1 = mods[0].weight, mods[0].bias
    e = e.to(dt)
    z = AG.linear(e, w1[:, 2 * C :], b1, dt)
    p_i = AG.linear(x_dst, w1[:, :C], None, dt)
    p_j = AG.linear(x_src, w1[:, C : 2 * C], None, dt)
    z = z + p_i.index_select(0, csr.dst32.long()) + p_j.index_select(0, csr.src32.long())
    h = AG.GeluFn.apply(z)
    i = 2
    while i < len(mods):
        act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight")
        h = lin(mods[i], h, dt, gelu=act)
        i += 2 if act else 1
    ln = m.layer_norm
    _plain_ln(ln)
    e_new, out = AG.graphconv_tail(h, ln.weight, ln.bias, e, csr, ln.eps)
    return out, e_new


def gnn_block(block, x_src: Tensor, x_dst: Tensor, edge_attr: Tensor, edge_index: Tensor, dt: torch.dtype, bipartite: bool,
              x_src_local: Optional[Tensor] = None):  # fmt: skip
    """GraphConvProcessorBlock / GraphConvMapperBlock forward (block.py:362-395, 441-479); returns ((src_new, dst_new), edges_new).
    ``x_src`` holds every source row the (local) edges name — on a model-parallel group the all-gathered rows (block.py:375, :451) —
    ``x_src_local`` this rank's source rows (the ones a forward mapper updates)."""
    if block.emb_edges is not None:
        edge_attr = mlp(block.emb_edges, edge_attr, dt)
    csr = Fn.csr_for(edge_index, x_src.shape[0], x_dst.shape[0])
    out, e_new = graph_conv(block.conv, x_src, x_dst, edge_attr, csr, dt)
    xd = x_dst.to(dt)
    dst_new = mlp(block.node_mlp, torch.cat([xd, out], 1), dt, residual=xd)
    src_new = x_src if x_src_local is None else x_src_local
    if bipartite and block.update_src_nodes:  # block.py:475 — the same node_mlp on cat[x_src, x_src]
        xs = src_new.to(dt)
        src_new = mlp(block.node_mlp, torch.cat([xs, xs], 1), dt, residual=xs)
    return (src_new, dst_new), e_new


# ------------------------------------------------------------------------------------------------------------
# GraphTransformer
# ------------------------------------------------------------------------------------------------------------
def gt_block(block, x_src: Optional[Tensor], x_dst: Tensor, edge_attr: Tensor, edge_index: Tensor, dt: torch.dtype, ln_src: Optional[nn.Module],
             cond=None, plan=None) -> Tensor:  # fmt: skip
    """GraphTransformerProcessorBlock (``x_src is None``: block.py:1219-1273) / GraphTransformerMapperBlock (block.py:963-1029) forward with the
    materialised edge projection (the reference's own formulation, block.py:623-635 + conv.py:103-147); returns the new dst rows."""
    if cond is not None:
        raise NotImplementedError("training with ConditionalLayerNorm conditioning: its backward is not implemented")
    A, H = block.attn_channels, block.num_heads
    xd_n = norm(block.layer_norm_attention, x_dst, dt)
    if x_src is None:
        buf = lin_cat([block.lin_query, block.lin_key, block.lin_value, block.lin_self], xd_n, dt)
        q, k, v, x_r = buf[:, :A], buf[:, A : 2 * A], buf[:, 2 * A : 3 * A], buf[:, 3 * A :]
        n_src = x_dst.shape[0]
    else:
        xs_n = norm(ln_src, x_src, dt)
        kv = lin_cat([block.lin_key, block.lin_value], xs_n, dt)
        qs = lin_cat([block.lin_query, block.lin_self], xd_n, dt)
        q, x_r, k, v = qs[:, :A], qs[:, A:], kv[:, :A], kv[:, A:]
        n_src = x_src.shape[0]
    if block.qk_norm:
        q = norm(block.q_norm, q, dt, groups=H)
        k = norm(block.k_norm, k, dt, groups=H)
    if plan is not None:
        # model-parallel "edges" strategy: k | v of the rows this rank owns, plus the halo rows its edges name on other ranks (differentiable
        # exchange); ``plan.edge_index`` addresses the compact table [own rows | halo rows] and local destination rows
        from ..distributed.graph import HaloExchangeFn

        kv_local = torch.cat([k, v], 1)
        table = torch.cat([kv_local, HaloExchangeFn.apply(kv_local, plan)], 0)
        k, v = table[:, :A], table[:, A:]
        edge_index, n_src = plan.edge_index, table.shape[0]
    ea = edge_attr
    if not isinstance(block.edge_pre_mlp, nn.Identity):
        ea = lin(block.edge_pre_mlp[0], ea, torch.float32, gelu=True)
    e = lin(block.lin_edge, ea, dt)
    csr = Fn.csr_for(edge_index, n_src, x_dst.shape[0])
    att = AG.gt_attention(q, k, v, e, csr, H)
    skip = x_dst.to(dt)
    o = lin(block.projection, att + x_r, dt) + skip
    return mlp(block.node_dst_mlp, o, dt, residual=o, pre_ln=block.layer_norm_mlp_dst)
