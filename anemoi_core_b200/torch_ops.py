"""Operator-level seam: the reference's fused attention custom op, served by libanemoi_b200.

The reference registers ``torch.library.custom_op("anemoi::graph_transformer_attention")`` with ``register_fake`` and
``register_autograd`` (triton/gt.py:390-447, 451-556) and calls it from ``GraphTransformerBaseBlock`` as
``self.conv(query, key, value, edges, csc, reverse)`` (layers/block.py:787-791).  Importing this module registers

    torch.ops.anemoi_b200.graph_transformer_attention(q, k, v, e, row, colptr, rowptr, edge_ids, edge_dst) -> (out, out_saved, m)

with the SAME signature, argument meaning and outputs (``out`` in ``q.dtype``, ``out_saved`` float32, ``m`` = float32 log-sum-exp per
(dst, head), zeros for rows without edges), a fake (meta) implementation for ``torch.compile`` tracing and an autograd formula
(``anemoi_b200_gt_attention_bwd``).  A reference maintainer who wants only the kernel swaps one import:

    from anemoi_core_b200.torch_ops import graph_transformer_attention_conv as graph_transformer_attention_conv

The namespace is ``anemoi_b200`` so that the reference's own op (namespace ``anemoi``) can stay registered next to it.
"""

from __future__ import annotations

import torch
from torch import Tensor

from . import ops
from ._ident import version

_CSR: dict = {}


def _csr_from_csc(row: Tensor, colptr: Tensor, n_src: int, rowptr: Tensor, edge_ids: Tensor, edge_dst: Tensor) -> ops.GraphCSR:
    """GraphCSR view of the reference's ``(row, colptr)`` / ``(rowptr, edge_ids, edge_dst)`` tensors (triton/utils.py:25-70), int32
    copies cached on the tensors' identity."""
    key = (row.data_ptr(), version(row), colptr.data_ptr(), version(colptr), tuple(row.shape), n_src)
    hit = _CSR.get(key)
    if hit is not None:
        return hit[1]
    n_dst, n_edges = colptr.numel() - 1, row.numel()
    colptr = colptr.contiguous()
    dst = edge_dst if edge_dst.numel() == n_edges else torch.repeat_interleave(torch.arange(n_dst, device=row.device), colptr[1:] - colptr[:-1])
    csr = ops.GraphCSR(n_src, n_dst, n_edges, colptr.long(), colptr.to(torch.int32), row.to(torch.int32).contiguous(), dst.to(torch.int32).contiguous())
    if rowptr.numel() == n_src + 1 and edge_ids.numel() == n_edges:
        csr.rev = (rowptr.to(torch.int32).contiguous(), edge_ids.to(torch.int32).contiguous())
    if len(_CSR) > 32:
        _CSR.clear()
    _CSR[key] = ((row, colptr), csr)
    return csr


@torch.library.custom_op("anemoi_b200::graph_transformer_attention", mutates_args=(), device_types="cuda")
def graph_transformer_attention(q: Tensor, k: Tensor, v: Tensor, e: Tensor, row: Tensor, colptr: Tensor, rowptr: Tensor, edge_ids: Tensor,
                                edge_dst: Tensor) -> tuple[Tensor, Tensor, Tensor]:  # fmt: skip
    n_dst, H, C = q.shape
    n_src = k.shape[0]
    csr = _csr_from_csc(row, colptr, n_src, rowptr, edge_ids, edge_dst)
    q2, k2, v2 = (t.contiguous().view(t.shape[0], H * C) for t in (q, k, v))
    e2 = e.contiguous().view(e.shape[0], H * C).to(q.dtype)
    m = torch.empty((n_dst, H), dtype=torch.float32, device=q.device)
    out = ops.gt_attention(q2, k2, v2, csr, H, e_proj=e2, lse=m).view(n_dst, H, C)
    return out, out.float() if out.dtype != torch.float32 else out.clone(), m


@graph_transformer_attention.register_fake
def _fake(q, k, v, e, row, colptr, rowptr, edge_ids, edge_dst):
    n_dst, H, C = q.shape
    return (torch.empty((n_dst, H, C), device=q.device, dtype=q.dtype), torch.empty((n_dst, H, C), device=q.device, dtype=torch.float32),
            torch.empty((n_dst, H), device=q.device, dtype=torch.float32))  # fmt: skip


def _setup_context(ctx, inputs, output):
    q, k, v, e, row, colptr, rowptr, edge_ids, edge_dst = inputs
    out, _, m = output
    ctx.save_for_backward(q, k, v, e, row, colptr, rowptr, edge_ids, edge_dst, out, m)


def _backward(ctx, d_out, d_out_saved, d_m):
    q, k, v, e, row, colptr, rowptr, edge_ids, edge_dst, out, m = ctx.saved_tensors
    n_dst, H, C = q.shape
    n_src = k.shape[0]
    csr = _csr_from_csc(row, colptr, n_src, rowptr, edge_ids, edge_dst)
    dt = q.dtype
    flat = lambda t: t.contiguous().view(t.shape[0], H * C).to(dt)  # noqa: E731
    dq, dk, dv, de = ops.gt_attention_bwd(flat(q), flat(k), flat(v), flat(e), flat(out), flat(d_out), m, csr, H)
    return dq.view(n_dst, H, C), dk.view(n_src, H, C), dv.view(n_src, H, C), de.view(e.shape[0], H, C).to(e.dtype), None, None, None, None, None


torch.library.register_autograd("anemoi_b200::graph_transformer_attention", _backward, setup_context=_setup_context)


def graph_transformer_attention_conv(query: Tensor, key: Tensor, value: Tensor, edges: Tensor, csc, reverse) -> Tensor:
    """Drop-in for the reference's ``graph_transformer_attention_conv`` (triton/gt.py, called at layers/block.py:787-791):
    ``csc = (row, colptr)``, ``reverse = (rowptr, edge_ids, edge_dst)`` from ``edge_index_to_csc(..., reverse=True)``."""
    row, colptr = csc
    rowptr, edge_ids, edge_dst = reverse
    return torch.ops.anemoi_b200.graph_transformer_attention(query, key, value, edges, row, colptr, rowptr, edge_ids, edge_dst)[0]
