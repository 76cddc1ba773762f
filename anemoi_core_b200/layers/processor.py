"""Processors: stacks of N message-passing blocks (reference: layers/processor.py:52-147, 319-626).

Drop-in for ``anemoi.models.layers.processor.{GNNProcessor, GraphTransformerProcessor}``: keyword-only constructors
with the reference's names, ``forward(x, batch_size, shard_info, edge_attr, edge_index, model_comm_group=None,
edges_are_dst_sorted=True)``, blocks held in ``self.proc`` (same ``state_dict`` keys).  Activation checkpointing
(``gradient_checkpointing``, one checkpoint per chunk of ``num_layers / num_chunks`` layers like the reference's ``run_layer_chunk``) applies
to the training path when ``ANEMOI_B200_ACT_CHECKPOINT=1`` (layers/_train.py:run_chunks; inert under ``no_grad``); CPU offload is refused.
"""

from __future__ import annotations

import os
from typing import Optional

import torch
from torch import Tensor
from torch import nn

from .. import ops
from ..distributed.graph import group_rank
from ..distributed.graph import group_size
from ..distributed.halo import tensor_ident
from ..distributed.khop_edges import build_graph_partition
from ..distributed.khop_edges import ensure_edges_are_dst_sorted
from ..distributed.shapes import GraphShardInfo
from . import _functional as Fn
from . import _reorder as RO
from . import _train as T
from .block import GraphConvProcessorBlock
from .block import GraphTransformerProcessorBlock
from .utils import compute_mlp_hidden_dim
from .utils import load_layer_kernels


class BaseProcessor(nn.Module):
    def __init__(
        self,
        *,
        num_layers: int,
        num_channels: int,
        num_chunks: int,
        cpu_offload: bool = False,
        gradient_checkpointing: bool = True,
        layer_kernels=None,
        **kwargs,
    ) -> None:
        super().__init__()
        if num_layers % num_chunks != 0:
            raise AssertionError(
                f"Number of processor layers ({num_layers}) has to be divisible by the number of processor chunks ({num_chunks})."
            )
        if cpu_offload:
            raise NotImplementedError("cpu_offload is a training-memory option of the reference; not part of the B200 forward path")
        self.num_layers = num_layers
        self.num_chunks = num_chunks
        self.chunk_size = num_layers // num_chunks
        self.num_channels = num_channels
        self.gradient_checkpointing = gradient_checkpointing
        self.layer_factory = load_layer_kernels(layer_kernels)

    def build_layers(self, layer_class, **layer_kwargs) -> None:
        self.proc = nn.ModuleList([layer_class(**layer_kwargs) for _ in range(self.num_layers)])


# sharded GNN processor: halo all-to-all of the source rows instead of the all-gather of x per layer.  Opt-in until it has run on NCCL
# (verified under Gloo, tests/test_sharded_forward_gloo.py); the GraphTransformer processor's halo exchange is the default (block.HALO_EXCHANGE).
GNN_HALO = os.environ.get("ANEMOI_B200_GNN_HALO", "0") == "1"
_SHARD_CACHE: dict = {}


def _shard_edges_by_dst(edge_attr: Tensor, edge_index: Tensor, n_dst: int, n_src: int, group, relabel_dst: bool = False, dst_splits=None):
    """Keep the edges into this rank's balanced dst range (reference ``shard_edges_1hop``, khop_edges.py:266-314).
    src ids stay global; dst ids stay global (GNN: the aggregate is computed over the full index range, block.py:375-391)
    or are relabelled to the local range (GraphTransformer).  The split is computed once per (graph, group) and cached.
    Returns (edge_attr view, edge_index, per-rank edge counts)."""
    world, rank = group_size(group), group_rank(group)
    key = (tensor_ident(edge_index), tuple(edge_index.shape), str(edge_index.device), n_dst, n_src, world, rank, relabel_dst,
           None if dst_splits is None else tuple(dst_splits))
    hit = _SHARD_CACHE.get(key)
    if hit is None:
        part = build_graph_partition(edge_index, world, (n_src, n_dst), dst_splits=dst_splits)
        e0, e1 = part.edge_range(rank)
        local = edge_index[:, e0:e1].clone()
        if relabel_dst:
            local[1] -= part.dst_range(rank)[0]
        hit = (edge_index, e0, e1, local.contiguous(), list(part.edge_splits))
        if len(_SHARD_CACHE) > 16:
            _SHARD_CACHE.clear()
        _SHARD_CACHE[key] = hit
    _, e0, e1, local, edge_sizes = hit
    return edge_attr[e0:e1], local, edge_sizes


_LOCAL_CACHE: dict = {}


def _localise_presharded_edges(edge_index: Tensor, dst_splits, group) -> Tensor:
    """Edges that arrive already cut to this rank's dst range (``StaticGraphProvider.get_edges(shard_edges=True)``, reference
    graph_provider.py:246-256) keep GLOBAL dst ids; the GraphTransformer kernels index local rows, so dst is relabelled once per
    (tensor, group) and the result cached on the tensor identity."""
    rank = group_rank(group)
    start = int(sum(dst_splits[:rank]))
    key = (tensor_ident(edge_index), tuple(edge_index.shape), str(edge_index.device), start)
    hit = _LOCAL_CACHE.get(key)
    if hit is None:
        local = edge_index.clone()
        local[1] -= start
        hit = (edge_index, local.contiguous())  # holding the source tensor keeps its data_ptr from being reused
        if len(_LOCAL_CACHE) > 16:
            _LOCAL_CACHE.clear()
        _LOCAL_CACHE[key] = hit
    return hit[1]


class GNNProcessor(BaseProcessor):
    """GraphConv processor (processor.py:319-455).  Layer 0 embeds the raw edge attributes; each layer hands its
    updated edge features to the next."""

    def __init__(
        self,
        *,
        num_channels: int,
        num_layers: int,
        num_chunks: int,
        mlp_extra_layers: int,
        edge_dim: int,
        mlp_hidden_ratio: float = 1.0,
        mlp_implementation: str = "mlp",
        cpu_offload: bool = False,
        layer_kernels=None,
        **kwargs,
    ) -> None:
        super().__init__(num_channels=num_channels, num_layers=num_layers, num_chunks=num_chunks, cpu_offload=cpu_offload,
                         layer_kernels=layer_kernels, **kwargs)  # fmt: skip
        build = dict(in_channels=num_channels, out_channels=num_channels, num_chunks=1, mlp_extra_layers=mlp_extra_layers,
                     mlp_hidden_ratio=mlp_hidden_ratio, mlp_implementation=mlp_implementation, layer_kernels=self.layer_factory)  # fmt: skip
        self.build_layers(GraphConvProcessorBlock, edge_dim=None, **build)
        self.proc[0] = GraphConvProcessorBlock(edge_dim=edge_dim, **build)

    def forward(
        self,
        x: Tensor,
        batch_size: int,
        shard_info: GraphShardInfo,
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        edges_are_dst_sorted: bool = True,
        *args,
        **kwargs,
    ) -> Tensor:
        n_nodes = sum(shard_info.nodes) if shard_info is not None and shard_info.nodes_are_sharded() else x.shape[0]
        if shard_info is None:
            shard_info = GraphShardInfo()
        if T.wants_grad(self, x, edge_attr):  # differentiable path: plain layer loop (layers/_train.py)
            world = group_size(model_comm_group)
            if not shard_info.edges_are_sharded():
                edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
                if world > 1:  # edges into this rank's rows, dst relabelled to the local range (sources: all-gathered per layer, global ids)
                    edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, n_nodes, n_nodes, model_comm_group, relabel_dst=True)
                    shard_info = GraphShardInfo(nodes=shard_info.nodes, edges=edge_sizes)
            elif world > 1:
                edge_index = _localise_presharded_edges(edge_index, shard_info.nodes, model_comm_group)
            group = model_comm_group if world > 1 else None

            def run_chunk(i: int, j: int, x: Tensor, edge_attr: Tensor):
                for block in self.proc[i:j]:
                    x, edge_attr = block(x, edge_attr, edge_index, shard_info, group)
                return x, edge_attr

            return T.run_chunks(self, run_chunk, x, edge_attr)[0]
        if not shard_info.edges_are_sharded():
            edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
            if group_size(model_comm_group) > 1:
                edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, n_nodes, n_nodes, model_comm_group)
                shard_info = GraphShardInfo(nodes=shard_info.nodes, edges=edge_sizes)
        halo_plan = None
        if GNN_HALO and group_size(model_comm_group) > 1:
            # halo exchange instead of the per-layer all-gather of x (block.py:375): needs local dst ids (the edges above keep global ones)
            from ..distributed.halo import halo_plan_for

            halo_plan = halo_plan_for(_localise_presharded_edges(edge_index, shard_info.nodes, model_comm_group), shard_info.nodes, model_comm_group)
        plan = RO.locality_plan(edge_index, n_nodes) if RO.ENABLED and group_size(model_comm_group) == 1 else None
        if plan is not None:  # run every layer in the locality order (layers/_reorder.py); identical results up to summation order
            x = ops.cast_pad(x, x.dtype, idx=plan.perm)
            edge_attr, edge_index = RO.permute_edge_attr(edge_attr, plan), plan.edge_index
        for block in self.proc:
            x, edge_attr = block(x, edge_attr, edge_index, shard_info, model_comm_group, halo_plan=halo_plan)
        return x if plan is None else ops.cast_pad(x, x.dtype, idx=plan.rank)


class GraphTransformerProcessor(BaseProcessor):
    """GraphTransformer processor (processor.py:458-626).  The raw edge attributes are shared by all layers; each
    layer owns its lin_edge, fused into its attention kernel."""

    def __init__(
        self,
        *,
        num_layers: int,
        num_channels: int,
        num_chunks: int,
        num_heads: int,
        mlp_hidden_ratio: float,
        edge_dim: int,
        attn_channels: Optional[int] = None,
        qk_norm: bool = False,
        mlp_implementation: str = "mlp",
        cpu_offload: bool = False,
        layer_kernels=None,
        shard_strategy: str = "edges",
        graph_attention_backend: str = "triton",
        edge_pre_mlp: bool = False,
        **kwargs,
    ) -> None:
        super().__init__(num_channels=num_channels, num_layers=num_layers, num_chunks=num_chunks, cpu_offload=cpu_offload,
                         layer_kernels=layer_kernels, **kwargs)  # fmt: skip
        if shard_strategy not in ("edges", "heads"):
            raise AssertionError(f"Invalid shard strategy '{shard_strategy}' for {self.__class__.__name__}. Supported strategies are 'edges' and 'heads'.")
        self.shard_strategy = shard_strategy
        self.build_layers(
            GraphTransformerProcessorBlock,
            in_channels=num_channels,
            hidden_dim=compute_mlp_hidden_dim(num_channels, mlp_hidden_ratio),
            out_channels=num_channels,
            attn_channels=attn_channels,
            num_heads=num_heads,
            layer_kernels=self.layer_factory,
            qk_norm=qk_norm,
            mlp_implementation=mlp_implementation,
            shard_strategy=shard_strategy,
            graph_attention_backend=graph_attention_backend,
            edge_dim=edge_dim,
            edge_pre_mlp=edge_pre_mlp,
        )

    def _tiled_attention(self, dt: torch.dtype) -> bool:
        """Will the blocks run the destination-tile attention kernel (the one that profits from the locality order)?"""
        from .block import ATTN_TILES

        blk = self.proc[0]
        return ATTN_TILES != "0" and blk._use_fold(dt) and ops.attention_tiles_supported(blk.attn_channels, blk.num_heads, dt, blk._fold_dims()[1])

    def forward(
        self,
        x: Tensor,
        batch_size: int,
        shard_info: GraphShardInfo,
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        edges_are_dst_sorted: bool = True,
        *args,
        **kwargs,
    ) -> Tensor:
        if shard_info is None:
            shard_info = GraphShardInfo()
        n_nodes = sum(shard_info.nodes) if shard_info.nodes_are_sharded() else x.shape[0]
        if T.wants_grad(self, x, edge_attr):  # differentiable path: plain layer loop (layers/_train.py)
            world = group_size(model_comm_group)
            heads = world > 1 and self.shard_strategy == "heads"  # attends over the FULL edge list for its heads: nothing to cut
            if heads and shard_info.edges_are_sharded():
                raise NotImplementedError("shard_strategy='heads' needs the full dst-sorted edge list (graph provider: get_edges(shard_edges=False))")
            if not shard_info.edges_are_sharded():
                edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
                if world > 1 and not heads:
                    edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, n_nodes, n_nodes, model_comm_group, relabel_dst=True)
                    shard_info = GraphShardInfo(nodes=shard_info.nodes, edges=edge_sizes)
            elif world > 1:
                edge_index = _localise_presharded_edges(edge_index, shard_info.nodes, model_comm_group)
            group, cond = model_comm_group if world > 1 else None, kwargs.get("cond")

            def run_chunk(i: int, j: int, x: Tensor, edge_attr: Tensor):
                for block in self.proc[i:j]:
                    x, _ = block(x, edge_attr, edge_index, shard_info, batch_size, n_nodes, group, cond=cond)
                return x, edge_attr

            return T.run_chunks(self, run_chunk, x, edge_attr)[0]
        if not shard_info.edges_are_sharded():
            edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
            if group_size(model_comm_group) > 1 and self.shard_strategy == "edges":
                # local dst rows only: dst relabelled to the local range (src ids stay global, sources are all-gathered per layer)
                edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, n_nodes, n_nodes, model_comm_group, relabel_dst=True)
                shard_info = GraphShardInfo(nodes=shard_info.nodes, edges=edge_sizes)
            # "heads": every rank attends over the FULL edge list for its heads (block._forward_heads), nothing to shard here
        elif group_size(model_comm_group) > 1 and self.shard_strategy == "heads":
            raise NotImplementedError("shard_strategy='heads' needs the full dst-sorted edge list (graph provider: get_edges(shard_edges=False))")
        elif group_size(model_comm_group) > 1:
            edge_index = _localise_presharded_edges(edge_index, shard_info.nodes, model_comm_group)
        cond = kwargs.get("cond")
        plan = None
        if group_size(model_comm_group) == 1 and RO.wanted(Fn.compute_dtype(x), self._tiled_attention(Fn.compute_dtype(x))):
            plan = RO.locality_plan(edge_index, n_nodes)  # cached per graph; the first call runs the host-side eigensolve
        if plan is not None:  # run every layer in the locality order (layers/_reorder.py); identical results up to summation order
            x = ops.cast_pad(x, x.dtype, idx=plan.perm)
            edge_attr, edge_index = RO.permute_edge_attr(edge_attr, plan), plan.edge_index
            if cond is not None:
                cond = ops.cast_pad(cond, cond.dtype, idx=plan.perm)
        shared_edges = None
        if all(isinstance(b.edge_pre_mlp, nn.Identity) for b in self.proc):
            shared_edges = self.proc[0].prepare_edges(edge_attr, Fn.compute_dtype(x))  # one padded fp32 copy for all layers
        for block in self.proc:
            x, _ = block(x, edge_attr, edge_index, shard_info, batch_size, n_nodes, model_comm_group, edge_attr_prepared=shared_edges,
                         cond=cond)  # fmt: skip
        return x if plan is None else ops.cast_pad(x, x.dtype, idx=plan.rank)
