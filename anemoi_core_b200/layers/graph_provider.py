"""Graph edge providers (reference: layers/graph_provider.py:37-291).

``StaticGraphProvider`` owns the dst-sorted edge list, the fixed edge attributes and the trainable edge tensor of one
sub-graph and hands ``(edge_attr, edge_index, edge_shard_sizes)`` to a mapper / processor on every forward.  Same
constructor, ``get_edges`` signature, ``edge_dim`` and ``state_dict`` keys (``trainable.trainable``,
``trainable_layout_version``) as the reference.

What differs is the cost: the reference re-concatenates ``[E, d_e]`` and re-expands ``edge_index`` per call (inside an
activation checkpoint), so downstream every forward sees new tensors and rebuilds its CSC (two sorts of E per layer,
block.py:779-782).  Here the expanded ``edge_index`` is built once per batch size, the attribute tensor once per version
of the trainable parameter, and both are returned by identity - the CSR plan (``layers/_functional.py:csr_for``), the
1-hop partition and the padded attribute rows are then cache hits for the whole run.
"""

from __future__ import annotations

from abc import ABC
from abc import abstractmethod
from typing import Optional

import torch
from torch import Tensor
from torch import nn

from .._ident import version
from ..distributed.graph import group_rank
from ..distributed.graph import group_size
from ..distributed.khop_edges import build_graph_partition
from ..distributed.khop_edges import sort_edge_index_by_dst
from .graph import TrainableTensor


def create_graph_provider(graph=None, edge_attributes: Optional[list[str]] = None, src_size: Optional[int] = None,
                          dst_size: Optional[int] = None, trainable_size: int = 0) -> "BaseGraphProvider":  # fmt: skip
    """``StaticGraphProvider`` when the sub-graph has edges, else ``NoOpGraphProvider`` (graph_provider.py:37-78)."""
    if graph:
        return StaticGraphProvider(graph=graph, edge_attributes=edge_attributes, src_size=src_size, dst_size=dst_size,
                                   trainable_size=trainable_size)  # fmt: skip
    return NoOpGraphProvider()


def normalize_projection_edges_name(edges_name) -> tuple[str, str, str]:
    """Only the explicit ``(src, "to", dst)`` triple is accepted (graph_provider.py:81-92)."""
    if not (isinstance(edges_name, (list, tuple)) and len(edges_name) == 3):
        raise ValueError(f"edges_name must be a (src, 'to', dst) triple, got {edges_name!r}")
    return tuple(edges_name)


class BaseGraphProvider(nn.Module, ABC):
    @abstractmethod
    def get_edges(self, batch_size=None, src_coords=None, dst_coords=None, model_comm_group=None, shard_edges: bool = True):
        """-> (edge_attr, edge_index, edge_shard_sizes)"""

    @property
    @abstractmethod
    def edge_dim(self) -> int: ...

    @property
    def is_sparse(self) -> bool:
        return False


def _edge_field(graph, name: str) -> Tensor:
    return graph[name] if not hasattr(graph, name) or isinstance(graph, dict) else getattr(graph, name)


class StaticGraphProvider(BaseGraphProvider):
    """Fixed edge structure + trainable edge features (graph_provider.py:145-291)."""

    _TRAINABLE_LAYOUT_VERSION = 1
    _TRAINABLE_LAYOUT_VERSION_KEY = "trainable_layout_version"

    def __init__(self, graph, edge_attributes: list[str], src_size: int, dst_size: int, trainable_size: int) -> None:
        super().__init__()
        assert graph, "StaticGraphProvider needs a valid graph to register edges."
        assert edge_attributes is not None, "Edge attributes must be provided"
        # sorted by destination once, here: everything downstream relies on contiguous dst runs (graph_provider.py:185-187)
        edge_index, perm = sort_edge_index_by_dst(_edge_field(graph, "edge_index"))
        edge_attr = torch.cat([_edge_field(graph, a) for a in edge_attributes], dim=1).index_select(0, perm)
        self.register_buffer("perm", perm, persistent=False)
        self.register_buffer("edge_attr", edge_attr, persistent=False)
        self.register_buffer("edge_index_base", edge_index.contiguous(), persistent=False)
        self.register_buffer("edge_inc", torch.tensor([[src_size], [dst_size]], dtype=torch.int64), persistent=False)
        self.register_buffer(self._TRAINABLE_LAYOUT_VERSION_KEY, torch.tensor(self._TRAINABLE_LAYOUT_VERSION, dtype=torch.int64), persistent=True)
        self.trainable = TrainableTensor(trainable_size=trainable_size, tensor_size=edge_attr.shape[0])
        self._edge_dim = edge_attr.shape[1] + trainable_size
        self._sizes = (int(src_size), int(dst_size))
        self._expanded: dict = {}  # batch_size -> (key, expanded edge_index)
        self._shards: dict = {}  # (batch_size, world, rank) -> (key, e0, e1, local edge_index, edge_splits)

    @property
    def edge_dim(self) -> int:
        return self._edge_dim

    def _expand_edges(self, edge_index: Tensor, edge_inc: Tensor, batch_size: int) -> Tensor:
        """``cat([edge_index + i * edge_inc for i in range(batch_size)], 1)`` (graph_provider.py:210-231), built once per batch size."""
        key = (edge_index.data_ptr(), version(edge_index), str(edge_index.device))
        hit = self._expanded.get(batch_size)
        if hit is None or hit[0] != key:
            out = edge_index if batch_size == 1 else torch.cat([edge_index + i * edge_inc for i in range(batch_size)], dim=1).contiguous()
            hit = (key, out)
            self._expanded[batch_size] = hit
        return hit[1]

    def _get_edges_impl(self, batch_size: int, shard_edges: bool, model_comm_group):
        edge_attr = self.trainable(self.edge_attr, batch_size)
        edge_index = self._expand_edges(self.edge_index_base, self.edge_inc, batch_size)
        if not shard_edges:
            return edge_attr, edge_index, None
        world = group_size(model_comm_group)
        if world <= 1:
            return edge_attr, edge_index, None  # shard_tensor with no group is the identity and edge_shard_sizes stays None (khop_edges.py:299-314)
        # 1-hop sharding of a dst-sorted list = contiguous edge ranges from the in-degrees (khop_edges.py:302-312)
        rank = group_rank(model_comm_group)
        skey = (batch_size, world, rank)
        key = (edge_index.data_ptr(), version(edge_index))
        hit = self._shards.get(skey)
        if hit is None or hit[0] != key:
            src_size, dst_size = self._sizes
            part = build_graph_partition(edge_index, world, (src_size * batch_size, dst_size * batch_size))
            e0, e1 = part.edge_range(rank)
            hit = (key, e0, e1, edge_index[:, e0:e1].contiguous(), list(part.edge_splits))
            self._shards[skey] = hit
        _, e0, e1, local_index, edge_splits = hit
        return edge_attr[e0:e1], local_index, edge_splits

    def get_edges(self, batch_size: int, src_coords: Optional[Tensor] = None, dst_coords: Optional[Tensor] = None, model_comm_group=None,
                  shard_edges: bool = True, act_checkpoint: bool = True):  # fmt: skip
        """``act_checkpoint`` is a training-memory device of the reference (recompute the concatenation in backward); the cached tensors make it moot."""
        return self._get_edges_impl(batch_size, shard_edges, model_comm_group)


class NoOpGraphProvider(BaseGraphProvider):
    """Edge-less architectures (graph_provider.py:294-338)."""

    @property
    def edge_dim(self) -> int:
        return 0

    def get_edges(self, batch_size=None, src_coords=None, dst_coords=None, model_comm_group=None, shard_edges: bool = True):
        return None, None, None
