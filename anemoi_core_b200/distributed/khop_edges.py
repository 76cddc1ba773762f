"""Integer path on the host side: stable dst-sort, O(1) dst-range partitions of a dst-sorted edge list, chunk
materialisation with src relabelling.  Reference: distributed/khop_edges.py:37-189, 412-500.  Pure index
bookkeeping in torch (CPU or CUDA tensors), bit-exact with the reference; it runs once per graph, not per step.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from .balanced_partition import get_balanced_partition_sizes


def sort_edge_index_by_dst(edge_index: Tensor) -> tuple[Tensor, Tensor]:
    """Stable sort by destination row; returns (sorted edge_index, permutation) (khop_edges.py:37-40)."""
    perm = torch.sort(edge_index[1], stable=True).indices
    return edge_index[:, perm], perm


def is_edge_index_dst_sorted(edge_index: Tensor) -> bool:
    dst = edge_index[1]
    return dst.numel() < 2 or bool((dst[1:] >= dst[:-1]).all())


def ensure_edges_are_dst_sorted(edge_attr: Tensor, edge_index: Tensor, edges_are_dst_sorted: bool = True) -> tuple[Tensor, Tensor]:
    """Sort (edge_attr, edge_index) by dst when the caller says they are not (khop_edges.py:242-263)."""
    if edges_are_dst_sorted:
        return edge_attr, edge_index
    edge_index, perm = sort_edge_index_by_dst(edge_index)
    return edge_attr[perm], edge_index


def drop_unconnected_src_nodes(n_src: int, edge_index: Tensor) -> tuple[Tensor, Tensor]:
    """Sorted unique connected src ids and the edge_index relabelled onto them (khop_edges.py:474-500)."""
    connected = torch.unique(edge_index[0])
    lut = torch.empty(n_src, dtype=torch.long, device=edge_index.device)
    lut[connected] = torch.arange(connected.numel(), device=edge_index.device)
    return connected, torch.stack([lut[edge_index[0]], edge_index[1]])


@dataclass(frozen=True)
class GraphPartition:
    """Contiguous dst ranges of a dst-sorted edge list and the matching edge ranges (khop_edges.py:51-151)."""

    dst_splits: tuple[int, ...]
    edge_splits: tuple[int, ...]
    num_nodes: tuple[int, int]

    @property
    def num_parts(self) -> int:
        return len(self.dst_splits)

    def dst_range(self, part: int) -> tuple[int, int]:
        start = sum(self.dst_splits[:part])
        return start, start + self.dst_splits[part]

    def edge_range(self, part: int) -> tuple[int, int]:
        start = sum(self.edge_splits[:part])
        return start, start + self.edge_splits[part]

    def materialise(self, part: int, edge_index: Tensor) -> tuple[tuple[int, int], tuple[int, int], Tensor, Tensor]:
        """(dst range, edge range, connected src ids, local edge_index with dst and src relabelled) of one part."""
        d0, d1 = self.dst_range(part)
        e0, e1 = self.edge_range(part)
        local = edge_index[:, e0:e1].clone()
        local[1] -= d0
        connected, local = drop_unconnected_src_nodes(self.num_nodes[0], local)
        return (d0, d1), (e0, e1), connected, local


def build_graph_partition(edge_index: Tensor, num_parts: int, num_nodes: tuple[int, int], dst_splits: Optional[list[int]] = None) -> GraphPartition:
    """Balanced dst splits (or the given ones) and per-part edge counts from the in-degrees (khop_edges.py:154-189)."""
    n_dst = num_nodes[1]
    if dst_splits is None:
        dst_splits = get_balanced_partition_sizes(n_dst, num_parts)
    if sum(dst_splits) != n_dst:
        raise ValueError("dst_splits do not sum to the number of destination nodes")
    degree = torch.bincount(edge_index[1], minlength=n_dst)
    bounds = torch.cumsum(torch.tensor([0] + list(dst_splits)), 0)
    csum = torch.cat([degree.new_zeros(1), torch.cumsum(degree, 0)]).cpu()
    edge_splits = [int(csum[bounds[i + 1]] - csum[bounds[i]]) for i in range(len(dst_splits))]
    return GraphPartition(tuple(int(s) for s in dst_splits), tuple(edge_splits), (int(num_nodes[0]), int(num_nodes[1])))
