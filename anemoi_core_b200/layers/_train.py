"""Differentiable forward of the blocks / processors / mappers: the same sm_100a kernels wired through
``anemoi_core_b200.autograd``, composed like the reference modules (layers/block.py:362-395, 441-479, 963-1029, 1219-1273,
layers/mlp.py:158-179) so that PyTorch autograd reaches every parameter, the inputs and the (trainable) edge attributes.

Taken whenever gradients are needed (``wants_grad``); the inference path under ``torch.no_grad()`` keeps its fusions (LayerNorm folded
into the GEMMs, lin_edge folded into attention, packed weights), which have no backward.  On a model-parallel group the "edges" strategy
trains through the autograd halves of the exchange (``distributed.graph.HaloExchangeFn`` / ``GatherRowsFn``; reference
distributed/graph.py:227-500): the forward collective is NCCL, the backward its transpose; parameter gradients are per-rank partial sums
(the trainer's gradient all-reduce over the model group completes them, as in the reference).  Gated feed-forward layers train through
``AG.glu_combine`` (one GEMM on the gate | value weights, as in the forward); a ConditionalLayerNorm is the affine-free LayerNorm kernel
followed by the two small conditioning Linears and an elementwise scale / shift in PyTorch (the per-row scale / bias ARE materialised in
training, as they are in the reference).  The heads (Ulysses) strategy trains through ``distributed.graph.AllToAllFn`` (the two exchanges of
block.py:689-759; backward = the same all-to-all with the split lists swapped).
"""

from __future__ import annotations

import os
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


# Activation checkpointing of the processors' training path (reference: layers/processor.py:129-147 ``run_layer_chunk`` under
# torch.utils.checkpoint, one checkpoint per chunk of num_layers / num_chunks layers, when ``gradient_checkpointing``): the chunk's forward is
# re-run in the backward instead of keeping its activations (~1.2 GB per 512-wide cfg2 layer).  The kernels are deterministic, so the recomputed
# forward is bit-identical; exchanges inside a chunk are re-run by every rank alike.  Opt-in: verified under the CPU stand-ins only
# (tests/test_sharded_training_gloo.py), it has not run on a GPU yet.
ACT_CHECKPOINT = os.environ.get("ANEMOI_B200_ACT_CHECKPOINT", "0") != "0"


def run_chunks(proc: nn.Module, run_chunk, *state):
    """``state = run_chunk(first_layer, last_layer, *state)`` over the processor's chunks, each under an activation checkpoint if asked for."""
    wrap = ACT_CHECKPOINT and getattr(proc, "gradient_checkpointing", False)
    for i in range(0, proc.num_layers, proc.chunk_size):
        j = min(i + proc.chunk_size, proc.num_layers)
        if wrap:
            from torch.utils.checkpoint import checkpoint

            state = checkpoint(run_chunk, i, j, *state, use_reentrant=False)
        else:
            state = run_chunk(i, j, *state)
    return state


def _single_gpu(group, what: str = "this module") -> None:
    from ..distributed.graph import group_size

    if group_size(group) > 1:
        raise NotImplementedError(f"training {what} on a model-parallel group is not implemented (the 'edges' strategy of the processors / mappers is)")


def _plain_ln(ln: nn.Module) -> None:
    from .normalization import _check_plain_layernorm

    _check_plain_layernorm(ln)


def norm(ln: nn.Module, x: Tensor, dt: torch.dtype, groups: int = 1, cond: Optional[Tensor] = None) -> Tensor:
    from .normalization import ConditionalLayerNorm

    if isinstance(ln, ConditionalLayerNorm):  # LN(x) * (1 + scale(cond)) + bias(cond)   (normalization.py:34-94)
        if cond is None:
            raise ValueError("ConditionalLayerNorm needs the conditioning tensor (cond=...)")
        xn = AG.layer_norm(x, None, None, ln.eps, torch.float32, groups)
        c = cond.reshape(-1, cond.shape[-1]).float()
        scale = AG.linear(c, ln.scale.weight, ln.scale.bias, torch.float32)
        shift = AG.linear(c, ln.bias.weight, ln.bias.bias, torch.float32)
        return (xn * (1.0 + scale) + shift).to(dt)
    _plain_ln(ln)
    return AG.layer_norm(x, ln.weight, getattr(ln, "bias", None), ln.eps, dt, groups)


FUSE_RES = os.environ.get("ANEMOI_B200_TRAIN_FUSE_RES", "1") != "0"  # residual adds in the GEMM epilogue (as in inference) instead of a PyTorch add


def lin(layer: nn.Module, x: Tensor, dt: torch.dtype, gelu: bool = False, residual: Optional[Tensor] = None) -> Tensor:
    if residual is not None and not FUSE_RES:
        y = AG.linear(x, layer.weight, getattr(layer, "bias", None), dt, gelu)
        return y + residual.to(y.dtype)
    return AG.linear(x, layer.weight, getattr(layer, "bias", None), dt, gelu, residual)


def lin_cat(layers, x: Tensor, dt: torch.dtype) -> Tensor:
    """One GEMM for several Linear containers on the same input (torch.cat is differentiable: the gradients split back by themselves)."""
    w = torch.cat([l.weight for l in layers], 0)
    bs = [getattr(l, "bias", None) for l in layers]
    b = None if all(x_ is None for x_ in bs) else torch.cat([b_ if b_ is not None else torch.zeros(l.weight.shape[0], device=w.device) for b_, l in zip(bs, layers)])
    return AG.linear(x, w, b, dt)


def mlp(m, x: Tensor, dt: torch.dtype, residual: Optional[Tensor] = None, pre_ln: Optional[nn.Module] = None, cond: Optional[Tensor] = None) -> Tensor:
    """``MLP.run`` with autograd: Linear(+GELU) / gated-layer chain, optional trailing LayerNorm, optional residual."""
    from .mlp import GatedMLPLayer

    if pre_ln is not None:
        x = norm(pre_ln, x, dt, cond=cond)
    mods = list(m.mlp)
    i = 0
    while i < len(mods):
        layer = mods[i]
        if isinstance(layer, GatedMLPLayer):  # act(gate_proj(x)) * value_proj(x): one GEMM on the concatenated weights, then the gating
            x = AG.glu_combine(lin_cat([layer.gate_proj, layer.value_proj], x, dt), layer.kind)
            i += 1
            continue
        act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight") and not isinstance(mods[i + 1], GatedMLPLayer)
        if residual is not None and not act and i + 1 == len(mods) and m.layer_norm is None:  # the chain ends in a plain Linear: + residual there
            return lin(layer, x, dt, residual=residual)
        x = lin(layer, x, dt, gelu=act)
        i += 2 if act else 1
    if m.layer_norm is not None:
        x = norm(m.layer_norm, x, dt, cond=cond)
    return x if residual is None else x + residual.to(x.dtype)


# ------------------------------------------------------------------------------------------------------------
# GraphConv (GNN)
# ------------------------------------------------------------------------------------------------------------
def graph_conv(conv, x_src: Tensor, x_dst: Tensor, e: Tensor, csr, dt: torch.dtype):
    """(out, e') of GraphConv (conv.py:66-81) with the split first layer: W1 [x_i; x_j; e] = (W1_i x_dst)[dst] + (W1_j x_src)[src] + W1_e e."""
    from .mlp import GatedMLPLayer

    m = conv.edge_mlp
    mods = list(m.mlp)
    if m.layer_norm is None:
        raise NotImplementedError("training GraphConv without the LayerNorm of its edge MLP")
    C = conv.in_channels
    gated0 = isinstance(mods[0], GatedMLPLayer)
    if gated0:  # gate | value rows of the first layer as one weight: the split over [x_i ; x_j ; e] applies to both
        firsts = [mods[0].gate_proj, mods[0].value_proj]
        w1 = torch.cat([l.weight for l in firsts], 0)
        bs = [l.bias for l in firsts]
        b1 = None if all(b is None for b in bs) else torch.cat([b if b is not None else torch.zeros(l.weight.shape[0], device=w1.device) for b, l in zip(bs, firsts)])
    else:
        w1, b1 = mods[0].weight, mods[0].bias
    e = e.to(dt)
    p_i = AG.linear(x_dst, w1[:, :C], None, dt)
    p_j = AG.linear(x_src, w1[:, C : 2 * C], None, dt)
    z = AG.edge_first_layer(e, w1[:, 2 * C :], b1, p_i, p_j, csr, dt)  # e W_e^T + b + p_i[dst] + p_j[src], gathers in the GEMM epilogue
    if gated0:
        h, i = AG.glu_combine(z, mods[0].kind), 1
    else:
        h, i = AG.GeluFn.apply(z), 2
    while i < len(mods):
        if isinstance(mods[i], GatedMLPLayer):
            h = AG.glu_combine(lin_cat([mods[i].gate_proj, mods[i].value_proj], h, dt), mods[i].kind)
            i += 1
            continue
        act = i + 1 < len(mods) and not hasattr(mods[i + 1], "weight") and not isinstance(mods[i + 1], GatedMLPLayer)
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
             cond=None, plan=None, heads=None) -> Tensor:  # fmt: skip
    """GraphTransformerProcessorBlock (``x_src is None``: block.py:1219-1273) / GraphTransformerMapperBlock (block.py:963-1029) forward with the
    materialised edge projection (the reference's own formulation, block.py:623-635 + conv.py:103-147); returns the new dst rows."""
    # ConditionalLayerNorm kernels: the processor block takes one conditioning tensor, the mapper block a (cond_src, cond_dst) pair
    cond_src, cond_dst = (cond if isinstance(cond, (tuple, list)) else (cond, cond)) if cond is not None else (None, None)
    A, H = block.attn_channels, block.num_heads
    xd_n = norm(block.layer_norm_attention if x_src is None else block.layer_norm_attention_dest, x_dst, dt, cond=cond_dst)
    if x_src is None:
        buf = lin_cat([block.lin_query, block.lin_key, block.lin_value, block.lin_self], xd_n, dt)
        q, k, v, x_r = buf.split(A, dim=1)  # split, not four slices: its backward is ONE cat instead of four zero-fill + add passes over [N, 4A]
        n_src = x_dst.shape[0]
    else:
        xs_n = norm(ln_src, x_src, dt, cond=cond_src)
        kv = lin_cat([block.lin_key, block.lin_value], xs_n, dt)
        qs = lin_cat([block.lin_query, block.lin_self], xd_n, dt)
        (q, x_r), (k, v) = qs.split(A, dim=1), kv.split(A, dim=1)
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
    if heads is not None:
        # model-parallel "heads" (Ulysses) strategy (block.py:689-759): (group, dst row counts, src row counts | None = sources replicated).
        # One all-to-all gives every rank ALL rows for its H / P heads, attention runs over the FULL edge list for those heads (the edge
        # projection's columns of those heads), a second all-to-all returns every rank its own rows for all heads.
        from ..distributed.graph import AllToAllFn
        from ..distributed.graph import group_rank
        from ..distributed.graph import group_size

        group, dst_sizes, src_sizes = heads
        P, me = group_size(group), group_rank(group)
        if H % P:
            raise ValueError(f"heads strategy: num_heads ({H}) must be divisible by the model group size ({P})")
        Hl, Ch = H // P, A // H
        n_l = x_dst.shape[0]

        def to_heads(t: Tensor, sizes: list) -> Tensor:  # [rows, H * Ch] -> all rows (global order) of my head group [sum(sizes), Hl * Ch]
            rows = t.shape[0]
            return AllToAllFn.apply(t.reshape(rows, P, Hl * Ch).permute(1, 0, 2).reshape(P * rows, Hl * Ch), [rows] * P, list(sizes), group)

        mine = slice(me * Hl * Ch, (me + 1) * Hl * Ch)
        q_h = to_heads(q, dst_sizes)
        k_h, v_h = (k[:, mine], v[:, mine]) if src_sizes is None else (to_heads(k, src_sizes), to_heads(v, src_sizes))
        csr = Fn.csr_for(edge_index, k_h.shape[0], sum(dst_sizes))
        att_h = AG.gt_attention(q_h, k_h.contiguous(), v_h.contiguous(), e[:, mine].contiguous(), csr, Hl)
        back = AllToAllFn.apply(att_h, list(dst_sizes), [n_l] * P, group)  # [head group (= source rank), local row, Hl * Ch]
        att = back.reshape(P, n_l, Hl * Ch).permute(1, 0, 2).reshape(n_l, A)
    else:
        csr = Fn.csr_for(edge_index, n_src, x_dst.shape[0])
        att = AG.gt_attention(q, k, v, e, csr, H)
    o = lin(block.projection, att + x_r, dt, residual=x_dst.to(dt))
    return mlp(block.node_dst_mlp, o, dt, residual=o, pre_ln=block.layer_norm_mlp_dst, cond=cond_dst)
