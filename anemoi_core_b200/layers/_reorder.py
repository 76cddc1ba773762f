"""Locality plan for a processor graph: an internal node order in which the sources of neighbouring destination nodes coincide.

Why: the attention / GraphConv kernels walk contiguous runs of destination nodes and gather their source rows.  The reference orders mesh
nodes by latitude only (``anemoi-graphs`` ``generate/utils.py:15-33``), so consecutive nodes sit at unrelated longitudes: on the ico-6
multi-scale mesh a tile of 16 consecutive destinations names 119 distinct sources for its 128 edges (reuse 1.07, median ``|src - dst|`` =
222 rows) and every gathered k | v row comes from L2.  Along a space-filling curve the same tile names 48 sources (reuse 2.65, median
distance 5): the gathers of a warp's node run hit L1 and a per-CTA source stage becomes possible.

How, from the topology alone (the processor API carries no coordinates): the three lowest non-trivial eigenvectors of the graph Laplacian of
a mesh on the sphere are the l = 1 spherical harmonics, i.e. the node positions up to a rotation (measured: 0.03 max coordinate error on
ico-6); the nodes are then sorted along a Hilbert curve on the faces of the enclosing cube.  One sparse shift-invert eigensolve on the host
per graph (4 s at 41 k nodes), cached on the edge_index tensor.  The processor permutes its input rows once, runs every layer in the new
order on a relabelled, dst-sorted edge list, and permutes the output back: results are identical up to summation order.

``ANEMOI_B200_REORDER``: "auto" (default) = on the bf16 path of the GraphTransformer processor, where the destination-tile attention
kernel (csrc/attention_tile.cu) turns the locality into 2.65x fewer gathered rows; "1" = always (both processors); "0" = never.
Measured (profiles/r2/): the warp-per-node kernel alone gains only 135 -> 130 us from the order (it is issue-bound, not L2-bound).
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from .._ident import version

MODE = os.environ.get("ANEMOI_B200_REORDER", "auto")
ENABLED = MODE == "1"  # unconditional (A/B switch); "auto" is decided per call by the processors (``wanted``)


def wanted(dt: torch.dtype, tiled_attention: bool) -> bool:
    """Should this forward run in the locality order?"""
    if MODE == "1":
        return True
    return MODE == "auto" and dt == torch.bfloat16 and tiled_attention


MIN_NODES = 2048  # below this the whole node tensor lives in L2 / L1 anyway


@dataclass
class ReorderPlan:
    perm: Tensor  # int32 [N]: new position -> old node id   (x_new = x_old[perm])
    rank: Tensor  # int32 [N]: old node id -> new position    (x_old = x_new[rank])
    edge_index: Tensor  # int64 [2, E]: relabelled, stable-sorted by the new dst
    edge_perm: Tensor  # int64 [E]: new edge position -> old edge position (edge_attr_new = edge_attr_old[edge_perm])


def _hilbert2(xi: np.ndarray, yi: np.ndarray, order: int) -> np.ndarray:
    """Index of the integer points (xi, yi) in [0, 2^order) along the 2-D Hilbert curve."""
    d = np.zeros_like(xi)
    x, y = xi.copy(), yi.copy()
    s = 1 << (order - 1)
    while s > 0:
        rx, ry = (x & s) > 0, (y & s) > 0
        d += s * s * ((3 * rx.astype(np.int64)) ^ ry.astype(np.int64))
        flip = ~ry & rx
        xr, yr = np.where(flip, s - 1 - x, x), np.where(flip, s - 1 - y, y)
        x, y = np.where(~ry, yr, xr) & (s - 1), np.where(~ry, xr, yr) & (s - 1)
        s >>= 1
    return d


def cube_hilbert_order(p: np.ndarray, order: int = 10) -> np.ndarray:
    """Sort 3-D directions ``p`` [N, 3] by (cube face, Hilbert index of the gnomonic projection on that face); returns new -> old."""
    n = p.shape[0]
    ax = np.argmax(np.abs(p), axis=1)
    major = p[np.arange(n), ax]
    face = ax * 2 + (major > 0)
    uv = np.zeros((n, 2))
    for a in range(3):
        m = ax == a
        o = [i for i in range(3) if i != a]
        den = np.maximum(np.abs(p[m, a]), 1e-12)
        uv[m, 0], uv[m, 1] = p[m, o[0]] / den, p[m, o[1]] / den
    q = np.clip(((uv + 1.0) * 0.5 * ((1 << order) - 1)).astype(np.int64), 0, (1 << order) - 1)
    return np.lexsort((_hilbert2(q[:, 0], q[:, 1], order), face))


def spectral_embedding(src: np.ndarray, dst: np.ndarray, n: int) -> Optional[np.ndarray]:
    """[N, 3] unit vectors from the 2nd..4th Laplacian eigenvectors of the symmetrised graph, or None if the solve fails."""
    try:
        import scipy.sparse as sp
        from scipy.sparse.linalg import eigsh

        a = sp.csr_matrix((np.ones(src.size), (dst, src)), shape=(n, n))
        a = ((a + a.T) > 0).astype(np.float64)
        a.setdiag(0)
        lap = sp.diags(np.asarray(a.sum(1)).ravel()) - a
        _, v = eigsh(lap.tocsc(), k=4, sigma=-1e-3, which="LM")
        emb = v[:, 1:4]
        nrm = np.linalg.norm(emb, axis=1, keepdims=True)
        if not np.isfinite(emb).all() or (nrm < 1e-12).any():
            return None
        return emb / nrm
    except Exception:  # noqa: BLE001 - any failure of the eigensolve just means "no reordering"
        return None


def source_reuse(edge_index: Tensor, tile: int = 16) -> float:
    """edges / distinct (dst tile, src) pairs: how often a tile of ``tile`` consecutive destinations re-reads a source row."""
    src, dst = edge_index[0].cpu().numpy().astype(np.int64), edge_index[1].cpu().numpy().astype(np.int64)
    if src.size == 0:
        return 1.0
    n = int(max(src.max(), dst.max())) + 1
    return float(src.size) / float(np.unique((dst // tile) * n + src).size)


_PLANS: dict = {}


def locality_plan(edge_index: Tensor, n_nodes: int, min_nodes: int = MIN_NODES, coords: Optional[np.ndarray] = None) -> Optional[ReorderPlan]:
    """Plan for a square (processor) graph, cached on the edge_index tensor; ``None`` when reordering is pointless or impossible.
    ``coords`` [N, 3] (unit vectors) skips the eigensolve."""
    key = (edge_index.data_ptr(), version(edge_index), tuple(edge_index.shape), str(edge_index.device), n_nodes)
    if key in _PLANS:
        return _PLANS[key][1]
    plan = None
    if n_nodes >= min_nodes and edge_index.shape[1] > 0:
        ei = edge_index.detach().cpu().numpy().astype(np.int64)
        emb = coords if coords is not None else spectral_embedding(ei[0], ei[1], n_nodes)
        if emb is not None:
            perm = cube_hilbert_order(np.asarray(emb, dtype=np.float64))
            rank = np.empty(n_nodes, np.int64)
            rank[perm] = np.arange(n_nodes)
            src, dst = rank[ei[0]], rank[ei[1]]
            eperm = np.argsort(dst, kind="stable")
            dev = edge_index.device

            def reuse(s_, d_):  # edges per distinct (16-row dst tile, src) pair
                return float(s_.size) / float(np.unique((d_ // 16) * n_nodes + s_).size)

            if coords is None and reuse(src, dst) < 1.2 * reuse(ei[0], ei[1]):
                emb = None  # the embedding found no locality worth two row permutations per forward (not a mesh-like graph)
        if emb is not None:
            plan = ReorderPlan(
                perm=torch.from_numpy(perm.astype(np.int32)).to(dev),
                rank=torch.from_numpy(rank.astype(np.int32)).to(dev),
                edge_index=torch.from_numpy(np.stack([src[eperm], dst[eperm]])).to(dev).contiguous(),
                edge_perm=torch.from_numpy(eperm).to(dev),
            )
    if len(_PLANS) > 16:
        _PLANS.clear()
    _PLANS[key] = (edge_index, plan)
    return plan


_ATTR_CACHE: dict = {}


def permute_edge_attr(edge_attr: Tensor, plan: ReorderPlan) -> Tensor:
    """``edge_attr[plan.edge_perm]``, cached on the attribute tensor (graph providers hand back the same tensor every step)."""
    key = (edge_attr.data_ptr(), version(edge_attr), tuple(edge_attr.shape), plan.edge_perm.data_ptr())
    hit = _ATTR_CACHE.get(key)
    if hit is None:
        hit = (edge_attr, edge_attr.index_select(0, plan.edge_perm).contiguous())
        if len(_ATTR_CACHE) > 16:
            _ATTR_CACHE.clear()
        _ATTR_CACHE[key] = hit
    return hit[1]
