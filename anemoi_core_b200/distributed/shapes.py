"""Shard metadata types that appear in every forward signature (reference: distributed/shapes.py:28-52).
Per-rank partition sizes along one tensor dimension, or None when the tensor is replicated."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional
from typing import Union

import torch.distributed as dist
from torch import Tensor

from .balanced_partition import get_balanced_partition_sizes

ShardSizes = Union[list[int], None]


@dataclass(frozen=True)
class GraphShardInfo:
    nodes: ShardSizes = None
    edges: ShardSizes = None

    def nodes_are_sharded(self) -> bool:
        return self.nodes is not None

    def edges_are_sharded(self) -> bool:
        return self.edges is not None


@dataclass(frozen=True)
class BipartiteGraphShardInfo:
    src_nodes: ShardSizes = None
    dst_nodes: ShardSizes = None
    edges: ShardSizes = None

    def src_is_sharded(self) -> bool:
        return self.src_nodes is not None

    def dst_is_sharded(self) -> bool:
        return self.dst_nodes is not None

    def edges_are_sharded(self) -> bool:
        return self.edges is not None


def get_shard_sizes(tensor: Tensor, dim: int, model_comm_group: Optional[dist.ProcessGroup] = None) -> ShardSizes:
    """Balanced per-rank sizes of ``tensor`` split along ``dim`` over the model group (shapes.py:55-62)."""
    if dim >= tensor.dim():
        raise ValueError(f"tensor has {tensor.dim()} dims, cannot split along {dim}")
    world = 1 if not model_comm_group else dist.get_world_size(group=model_comm_group)
    return get_balanced_partition_sizes(tensor.shape[dim], world)
