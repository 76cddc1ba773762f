"""Synthetic O96-like grid / icosahedral multi-scale mesh graphs for the benchmark and the parity tests.

Shapes follow SURVEY.md §8: an octahedral reduced-Gaussian-like grid (O96 = 4*96*105 = 40 320 points),
an icosphere of refinement level r (10*4^r + 2 vertices) with multi-scale edges = union of the icosphere
edges of levels 1..r (60*4^l directed edges per level; reference
``graphs/.../generate/tri_icosahedron.py:155-199`` + ``edges/builders/multi_scale.py:67-69``), encoder
edges by cut-off radius 0.6 x mesh reference distance capped at 64 neighbours
(``edges/builders/cutoff.py:123-151``), decoder edges by 3 nearest mesh nodes
(``edges/builders/knn.py``), edge attributes = [length, 2-D direction] (unit-std normalised) + 8
"trainable" columns.  All edge lists are returned stable-sorted by destination, int64, like
``StaticGraphProvider`` hands them to the model (``layers/graph_provider.py:185``).

Pure numpy/scipy, CPU, deterministic for a given seed.  This is input generation, not the hot path.
"""

from __future__ import annotations

import numpy as np
import torch

_ICO_CACHE: dict = {}


def _icosahedron():
    phi = (1.0 + 5.0**0.5) / 2.0
    v = np.array(
        [[-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0], [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
         [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], dtype=np.float64)  # fmt: skip
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(
        [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
         [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
         [8, 6, 7], [9, 8, 1]], dtype=np.int64)  # fmt: skip
    return v, f


def _subdivide(v: np.ndarray, f: np.ndarray):
    """One 1->4 triangle split; old vertices keep their index, midpoints are appended."""
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    e_sorted = np.sort(e, axis=1)
    uniq, inv = np.unique(e_sorted, axis=0, return_inverse=True)
    mid = v[uniq[:, 0]] + v[uniq[:, 1]]
    mid /= np.linalg.norm(mid, axis=1, keepdims=True)
    nv = v.shape[0]
    m = nv + inv.reshape(3, -1).T  # [F, 3]: midpoints of (01, 12, 20)
    a, b, c = f[:, 0], f[:, 1], f[:, 2]
    ab, bc, ca = m[:, 0], m[:, 1], m[:, 2]
    f_new = np.concatenate(
        [np.stack([a, ab, ca], 1), np.stack([b, bc, ab], 1), np.stack([c, ca, bc], 1), np.stack([ab, bc, ca], 1)], axis=0
    )
    return np.concatenate([v, mid], axis=0), f_new


def _directed_edges(f: np.ndarray) -> np.ndarray:
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    e = np.concatenate([e, e[:, ::-1]], axis=0)
    return np.unique(e, axis=0)  # 60 * 4^l directed edges


def icosphere_multiscale(level: int):
    """Vertices [10*4^level+2, 3] (lat-sorted like the reference's node ordering) and the union of the
    directed icosphere edges of refinement levels 1..level, as (src, dst) int64 [2, E]."""
    if level in _ICO_CACHE:
        return _ICO_CACHE[level]
    v, f = _icosahedron()
    edges = []
    for _ in range(1, level + 1):
        v, f = _subdivide(v, f)
        edges.append(_directed_edges(f))
    e = np.concatenate(edges, axis=0)
    # reference orders nodes by latitude then longitude (graphs/.../generate/utils.py:15-33)
    lat = np.arcsin(np.clip(v[:, 2], -1, 1))
    lon = np.arctan2(v[:, 1], v[:, 0])
    order = np.lexsort((lon, -lat))
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    v = v[order]
    e = rank[e]
    _ICO_CACHE[level] = (v, e.T.copy())
    return _ICO_CACHE[level]


def octahedral_grid(n: int) -> np.ndarray:
    """Octahedral reduced Gaussian grid O<n>: 2n latitudes, 20 + 4*i points on the i-th row from each pole
    (O96: 40 320 points).  Returns unit vectors [P, 3], north to south."""
    x, _ = np.polynomial.legendre.leggauss(2 * n)
    lats = np.arcsin(x[::-1])
    pts = []
    for i, la in enumerate(lats):
        k = i if i < n else 2 * n - 1 - i
        npts = 20 + 4 * k
        lo = np.arange(npts) * (2 * np.pi / npts)
        pts.append(np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.full(npts, np.sin(la))], 1))
    return np.concatenate(pts, axis=0)


def reduced_grid(n_points: int, n_lat: int) -> np.ndarray:
    """A reduced-Gaussian-like grid with exactly ``n_points`` points on ``n_lat`` latitudes, row lengths
    proportional to cos(lat) (stand-in for N320 = 542 080 points on 640 latitudes)."""
    x, _ = np.polynomial.legendre.leggauss(n_lat)
    lats = np.arcsin(x[::-1])
    w = np.cos(lats)
    counts = np.maximum(4, np.floor(w / w.sum() * n_points).astype(np.int64))
    diff = n_points - int(counts.sum())
    order = np.argsort(-w)
    i = 0
    while diff != 0:
        counts[order[i % n_lat]] += 1 if diff > 0 else -1
        diff += -1 if diff > 0 else 1
        i += 1
    pts = []
    for la, npts in zip(lats, counts):
        lo = np.arange(npts) * (2 * np.pi / npts)
        pts.append(np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.full(npts, np.sin(la))], 1))
    return np.concatenate(pts, axis=0)


def _sort_by_dst(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    perm = np.argsort(dst, kind="stable")
    return np.stack([src[perm], dst[perm]], axis=0).astype(np.int64)


def cutoff_edges(src_xyz: np.ndarray, dst_xyz: np.ndarray, factor: float = 0.6, max_nbrs: int = 64) -> np.ndarray:
    """All src within ``factor`` x (max nearest-neighbour distance among dst) of each dst, nearest first, capped."""
    from scipy.spatial import cKDTree

    tree_d = cKDTree(dst_xyz)
    d2, _ = tree_d.query(dst_xyz, k=2)
    radius = float(d2[:, 1].max()) * factor
    tree_s = cKDTree(src_xyz)
    nbrs = tree_s.query_ball_point(dst_xyz, r=radius, return_sorted=False)
    src, dst = [], []
    for d, lst in enumerate(nbrs):
        if not lst:
            continue
        lst = np.asarray(lst, dtype=np.int64)
        if lst.size > max_nbrs:
            dist = np.linalg.norm(src_xyz[lst] - dst_xyz[d], axis=1)
            lst = lst[np.argsort(dist, kind="stable")[:max_nbrs]]
        lst = np.sort(lst)
        src.append(lst)
        dst.append(np.full(lst.size, d, dtype=np.int64))
    return _sort_by_dst(np.concatenate(src), np.concatenate(dst))


def knn_edges(src_xyz: np.ndarray, dst_xyz: np.ndarray, k: int = 3) -> np.ndarray:
    from scipy.spatial import cKDTree

    _, idx = cKDTree(src_xyz).query(dst_xyz, k=k)
    idx = np.sort(idx.reshape(-1, k), axis=1)
    dst = np.repeat(np.arange(dst_xyz.shape[0], dtype=np.int64), k)
    return _sort_by_dst(idx.reshape(-1).astype(np.int64), dst)


def edge_attributes(src_xyz, dst_xyz, edge_index: np.ndarray, n_trainable: int, rng: np.random.Generator) -> np.ndarray:
    """[E, 3 + n_trainable] float32: chord length, 2-D direction of src seen from dst (local east/north
    tangent frame), each geometric column normalised to unit std; trainable columns ~ N(0, 0.1)."""
    s, d = src_xyz[edge_index[0]], dst_xyz[edge_index[1]]
    diff = s - d
    length = np.linalg.norm(diff, axis=1)
    lat = np.arcsin(np.clip(d[:, 2], -1, 1))
    lon = np.arctan2(d[:, 1], d[:, 0])
    east = np.stack([-np.sin(lon), np.cos(lon), np.zeros_like(lon)], 1)
    north = np.stack([-np.sin(lat) * np.cos(lon), -np.sin(lat) * np.sin(lon), np.cos(lat)], 1)
    de, dn = (diff * east).sum(1), (diff * north).sum(1)
    nrm = np.maximum(np.sqrt(de**2 + dn**2), 1e-12)
    geo = np.stack([length, de / nrm, dn / nrm], 1)
    geo = geo / np.maximum(geo.std(axis=0, keepdims=True), 1e-12)
    train = rng.normal(0.0, 0.1, size=(edge_index.shape[1], n_trainable))
    return np.concatenate([geo, train], axis=1).astype(np.float32)


def build_graph(grid: str = "o96", mesh_level: int = 6, n_trainable: int = 8, seed: int = 42) -> dict:
    """Encoder / processor / decoder graphs as torch tensors (CPU).  Keys: n_grid, n_mesh, edge_dim,
    {enc,proc,dec}_index [2,E] int64 dst-sorted, {enc,proc,dec}_attr [E, edge_dim] fp32."""
    rng = np.random.default_rng(seed)
    if grid == "o96":
        g = octahedral_grid(96)
    elif grid == "o32":
        g = octahedral_grid(32)
    elif grid == "n320":
        g = reduced_grid(542_080, 640)
    elif grid == "o1280":
        g = octahedral_grid(1280)
    else:
        raise ValueError(f"unknown grid {grid!r}")
    m, proc = icosphere_multiscale(mesh_level)
    proc = _sort_by_dst(proc[0], proc[1])
    enc = cutoff_edges(g, m, 0.6, 64)
    dec = knn_edges(m, g, 3)
    out = {"n_grid": int(g.shape[0]), "n_mesh": int(m.shape[0]), "edge_dim": 3 + n_trainable}
    for name, ei, (s, d) in (("enc", enc, (g, m)), ("proc", proc, (m, m)), ("dec", dec, (m, g))):
        out[name + "_index"] = torch.from_numpy(ei)
        out[name + "_attr"] = torch.from_numpy(edge_attributes(s, d, ei, n_trainable, rng))
    return out


def random_graph(n_src: int, n_dst: int, n_edges: int, edge_dim: int, seed: int = 42, sort: bool = True) -> dict:
    """cfg1-style random graph: src, dst ~ U, stable dst-sort, edge_attr ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, n_src, (n_edges,), generator=g), torch.randint(0, n_dst, (n_edges,), generator=g)])
    if sort:
        ei = ei[:, torch.sort(ei[1], stable=True)[1]]
    return {"index": ei, "attr": torch.randn(n_edges, edge_dim, generator=g)}
