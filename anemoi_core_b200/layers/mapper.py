"""Grid<->mesh mappers (reference: layers/mapper.py:51-1087).

Drop-in for ``anemoi.models.layers.mapper.{GraphTransformerForwardMapper, GraphTransformerBackwardMapper,
GNNForwardMapper, GNNBackwardMapper}``: same keyword-only constructors, ``forward(x=(x_src, x_dst), batch_size,
shard_info, edge_attr, edge_index, model_comm_group=None, keep_x_dst_sharded=..., edges_are_dst_sorted=True)``,
same return values (Forward: ``(x_src, x_dst_out)`` — the GraphTransformer one returns the *unembedded* ``x[0]``,
mapper.py:597, the GNN one the embedded+updated src, mapper.py:834; Backward: ``x_dst_out``), same ``state_dict`` keys.

``num_chunks`` (and ANEMOI_INFERENCE_NUM_CHUNKS_MAPPER) exist in the reference only to bound the [E, C] temporaries of
the unfused PyG/Triton path (mapper.py:289-295, 365-381).  The fused kernels never materialise per-edge channel
tensors for the GraphTransformer path, so the dst-chunk loop is not needed; the argument is accepted and results are
identical to any chunk count (that invariance is what the reference tests, test_graphtransformer_mapper.py:199-256).
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch import nn

from ..distributed.graph import group_size
from ..distributed.khop_edges import ensure_edges_are_dst_sorted
from ..distributed.shapes import BipartiteGraphShardInfo
from . import _functional as Fn
from . import _train as T
from .block import GraphConvMapperBlock
from .block import GraphTransformerMapperBlock
from .mlp import MLP
from .utils import compute_mlp_hidden_dim
from .utils import load_layer_kernels

PairTensor = tuple[Tensor, Tensor]


class BaseMapper(Fn.PackOwner):
    def __init__(
        self,
        *,
        in_channels_src: int,
        in_channels_dst: int,
        hidden_dim: int,
        out_channels_dst: Optional[int] = None,
        cpu_offload: bool = False,
        gradient_checkpointing: bool = True,
        layer_kernels=None,
        **kwargs,
    ) -> None:
        super().__init__()
        if cpu_offload:
            raise NotImplementedError("cpu_offload is a training-memory option of the reference; not part of the B200 forward path")
        self.in_channels_src = in_channels_src
        self.in_channels_dst = in_channels_dst
        self.hidden_dim = hidden_dim
        self.out_channels_dst = out_channels_dst
        self.gradient_checkpointing = gradient_checkpointing
        self.layer_factory = load_layer_kernels(layer_kernels)
        self._pack = Fn.WeightPack()


# ------------------------------------------------------------------------------------------------------------
# GraphTransformer mappers
# ------------------------------------------------------------------------------------------------------------
class GraphTransformerBaseMapper(BaseMapper):
    def __init__(
        self,
        *,
        in_channels_src: int,
        in_channels_dst: int,
        hidden_dim: int,
        out_channels_dst: Optional[int] = None,
        num_chunks: int = 1,
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
        super().__init__(in_channels_src=in_channels_src, in_channels_dst=in_channels_dst, hidden_dim=hidden_dim,
                         out_channels_dst=out_channels_dst, cpu_offload=cpu_offload, layer_kernels=layer_kernels, **kwargs)  # fmt: skip
        if shard_strategy not in ("heads", "edges"):
            raise AssertionError(f"Invalid shard strategy '{shard_strategy}' for {self.__class__.__name__}. Supported strategies are 'heads' and 'edges'.")
        self.num_chunks = num_chunks
        self.shard_strategy = shard_strategy
        self.proc = GraphTransformerMapperBlock(
            in_channels=hidden_dim,
            hidden_dim=compute_mlp_hidden_dim(hidden_dim, mlp_hidden_ratio),
            out_channels=hidden_dim,
            attn_channels=attn_channels,
            num_heads=num_heads,
            edge_dim=edge_dim,
            qk_norm=qk_norm,
            mlp_implementation=mlp_implementation,
            layer_kernels=self.layer_factory,
            shard_strategy=shard_strategy,
            graph_attention_backend=graph_attention_backend,
            edge_pre_mlp=edge_pre_mlp,
        )
        self.emb_nodes_dst = self.layer_factory.Linear(self.in_channels_dst, self.hidden_dim)

    def pre_process(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        raise NotImplementedError

    def post_process(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return x_dst

    def _run(self, x: PairTensor, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted,
             keep_x_dst_sharded: bool = True, cond=None) -> Tensor:  # fmt: skip
        """Single GPU, or dst-range sharded over ``model_comm_group`` (reference "edges" strategy, mapper.py:248-386):
        ``x[1]`` is then this rank's slice of the destination rows (``shard_info.dst_nodes``), ``x[0]`` either the full source
        tensor (``shard_info.src_nodes is None``) or this rank's slice (its k | v rows are all-gathered inside the block);
        the full dst-sorted edge list is cut to the edges into the local rows (cached per graph and group)."""
        edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
        world = group_size(model_comm_group)
        train = T.wants_grad(self, x[0], x[1], edge_attr)  # differentiable path (layers/_train.py): same call sequence, same sharding
        if world > 1:
            if shard_info is None or not shard_info.dst_is_sharded():
                raise ValueError("sharded mapper: shard_info.dst_nodes (per-rank destination row counts) is required")
            from .processor import _localise_presharded_edges
            from .processor import _shard_edges_by_dst

            if self.shard_strategy == "heads":
                if shard_info.edges_are_sharded():
                    raise NotImplementedError("shard_strategy='heads' needs the full dst-sorted edge list (graph provider: get_edges(shard_edges=False))")
                # every rank attends over the FULL edge list for its heads (block._heads_attention): nothing to cut
            elif shard_info.edges_are_sharded():
                # the graph provider already cut the list to the edges into our rows (global dst ids): relabel dst only
                edge_index = _localise_presharded_edges(edge_index, shard_info.dst_nodes, model_comm_group)
            else:
                n_src = sum(shard_info.src_nodes) if shard_info.src_is_sharded() else x[0].shape[0]
                edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, sum(shard_info.dst_nodes), n_src, model_comm_group,
                                                                        relabel_dst=True, dst_splits=shard_info.dst_nodes)  # fmt: skip
                shard_info = BipartiteGraphShardInfo(src_nodes=shard_info.src_nodes, dst_nodes=shard_info.dst_nodes, edges=edge_sizes)
        dt = Fn.compute_dtype(*x)
        x_src, x_dst = (self.pre_process_train if train else self.pre_process)(x, dt)
        (_, x_dst_out), _ = self.proc((x_src, x_dst), edge_attr, edge_index, shard_info, batch_size, (x_src.shape[0], x_dst.shape[0]),
                                      model_comm_group if world > 1 else None, cond=cond)  # fmt: skip
        out = (self.post_process_train if train else self.post_process)(x_dst_out, dt)
        if world > 1 and not keep_x_dst_sharded:
            from ..distributed.graph import gather_rows
            from ..distributed.graph import gather_rows_grad

            out = (gather_rows_grad if train else gather_rows)(out, shard_info.dst_nodes, model_comm_group)
        return out


class GraphTransformerForwardMapper(GraphTransformerBaseMapper):
    """data -> hidden (mapper.py:480-597)."""

    def __init__(self, *, out_channels_dst: Optional[int] = None, **kwargs) -> None:
        if out_channels_dst is not None:
            raise AssertionError("GraphTransformerForwardMapper does not support out_channels_dst.")
        super().__init__(out_channels_dst=None, **kwargs)
        self.emb_nodes_src = self.layer_factory.Linear(self.in_channels_src, self.hidden_dim)

    def pre_process(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        x_src, x_dst = x
        return Fn.fused_linear(self._pack, x_src, [self.emb_nodes_src], dt), Fn.fused_linear(self._pack, x_dst, [self.emb_nodes_dst], dt)

    def pre_process_train(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return T.lin(self.emb_nodes_src, x[0], dt), T.lin(self.emb_nodes_dst, x[1], dt)

    def post_process_train(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return x_dst

    def forward(
        self,
        x: PairTensor,
        batch_size: int,
        shard_info: Optional[BipartiteGraphShardInfo],
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        keep_x_dst_sharded: bool = True,
        edges_are_dst_sorted: bool = True,
        **kwargs,
    ) -> PairTensor:
        return x[0], self._run(x, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted, keep_x_dst_sharded, cond=kwargs.get("cond"))


class GraphTransformerBackwardMapper(GraphTransformerBaseMapper):
    """hidden -> data (mapper.py:600-704): dst embedded by ``emb_nodes_dst``; extractor = LayerNorm + Linear."""

    def __init__(self, *, initialise_data_extractor_zero: bool = False, **kwargs) -> None:
        super().__init__(**kwargs)
        k = self.layer_factory
        self.node_data_extractor = nn.Sequential(k.LayerNorm(normalized_shape=self.hidden_dim), k.Linear(self.hidden_dim, self.out_channels_dst))
        if initialise_data_extractor_zero:
            for module in self.node_data_extractor.modules():
                if isinstance(module, nn.Linear):
                    nn.init.constant_(module.weight, 0.0)
                    if module.bias is not None:
                        nn.init.constant_(module.bias, 0.0)

    def pre_process(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        x_src, x_dst = x
        return x_src, Fn.fused_linear(self._pack, x_dst, [self.emb_nodes_dst], dt)

    def post_process(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        lin = self.node_data_extractor[1]
        return Fn.ln_linear(self._pack, x_dst, self.node_data_extractor[0], ("extractor",), Fn.linear_sources([lin]), lambda: Fn.cat_linear32([lin]), dt)

    def pre_process_train(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return x[0], T.lin(self.emb_nodes_dst, x[1], dt)

    def post_process_train(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return T.lin(self.node_data_extractor[1], T.norm(self.node_data_extractor[0], x_dst, dt), dt)

    def forward(
        self,
        x: PairTensor,
        batch_size: int,
        shard_info: Optional[BipartiteGraphShardInfo],
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        keep_x_dst_sharded: bool = False,
        edges_are_dst_sorted: bool = True,
        **kwargs,
    ) -> Tensor:
        return self._run(x, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted, keep_x_dst_sharded, cond=kwargs.get("cond"))


# ------------------------------------------------------------------------------------------------------------
# GNN (GraphConv) mappers
# ------------------------------------------------------------------------------------------------------------
class GNNBaseMapper(BaseMapper):
    def __init__(
        self,
        *,
        in_channels_src: int,
        in_channels_dst: int,
        hidden_dim: int,
        out_channels_dst: Optional[int] = None,
        num_chunks: int = 1,
        mlp_extra_layers: int,
        edge_dim: int,
        mlp_hidden_ratio: float = 1.0,
        mlp_implementation: str = "mlp",
        cpu_offload: bool = False,
        layer_kernels=None,
        **kwargs,
    ) -> None:
        super().__init__(in_channels_src=in_channels_src, in_channels_dst=in_channels_dst, hidden_dim=hidden_dim,
                         out_channels_dst=out_channels_dst, cpu_offload=cpu_offload, layer_kernels=layer_kernels, **kwargs)  # fmt: skip
        self.num_chunks = num_chunks
        self._mlp_kw = dict(layer_kernels=self.layer_factory, n_extra_layers=mlp_extra_layers + 1, mlp_implementation=mlp_implementation)
        self._mlp_hidden = compute_mlp_hidden_dim(hidden_dim, mlp_hidden_ratio)
        self.emb_edges = MLP(edge_dim, self._mlp_hidden, hidden_dim, **self._mlp_kw)
        self._block_kw = dict(in_channels=hidden_dim, out_channels=hidden_dim, num_chunks=num_chunks, mlp_extra_layers=mlp_extra_layers,
                              mlp_hidden_ratio=mlp_hidden_ratio, mlp_implementation=mlp_implementation, layer_kernels=self.layer_factory)  # fmt: skip

    def pre_process(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return x

    def post_process(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return x_dst

    def pre_process_train(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return x

    def post_process_train(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return x_dst

    def _run(self, x: PairTensor, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted,
             keep_x_dst_sharded: bool = False) -> PairTensor:  # fmt: skip
        """Single GPU, or sharded over ``model_comm_group`` like the reference (mapper.py:760-835): src and dst rows are sharded (a replicated
        input is cut to this rank's balanced slice, ``ensure_sharded``), the edges are the ones into the local dst rows, the block all-gathers
        the embedded src rows (``sync_tensor``, block.py:451); returns the LOCAL src shard and the dst rows (gathered unless
        ``keep_x_dst_sharded``)."""
        edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, edges_are_dst_sorted)
        world = group_size(model_comm_group)
        x_src, x_dst = x
        train = T.wants_grad(self, x_src, x_dst, edge_attr)  # differentiable path (layers/_train.py): same call sequence, same sharding
        if world > 1:
            from ..distributed.balanced_partition import get_balanced_partition_sizes
            from ..distributed.graph import shard_rows
            from .processor import _localise_presharded_edges
            from .processor import _shard_edges_by_dst

            if shard_info is None:
                shard_info = BipartiteGraphShardInfo()
            src_sizes, dst_sizes = shard_info.src_nodes, shard_info.dst_nodes
            if src_sizes is None:  # replicated input: keep this rank's balanced slice
                src_sizes = get_balanced_partition_sizes(x_src.shape[0], world)
                x_src = shard_rows(x_src, src_sizes, model_comm_group)
            if dst_sizes is None:
                dst_sizes = get_balanced_partition_sizes(x_dst.shape[0], world)
                x_dst = shard_rows(x_dst, dst_sizes, model_comm_group)
            if shard_info.edges_are_sharded():
                edge_index, edge_sizes = _localise_presharded_edges(edge_index, dst_sizes, model_comm_group), shard_info.edges
            else:
                edge_attr, edge_index, edge_sizes = _shard_edges_by_dst(edge_attr, edge_index, sum(dst_sizes), sum(src_sizes), model_comm_group,
                                                                        relabel_dst=True, dst_splits=dst_sizes)  # fmt: skip
            shard_info = BipartiteGraphShardInfo(src_nodes=src_sizes, dst_nodes=dst_sizes, edges=edge_sizes)
        dt = Fn.compute_dtype(x_src, x_dst, edge_attr)
        e = T.mlp(self.emb_edges, edge_attr, dt) if train else self.emb_edges.run(edge_attr, dt)
        x_src, x_dst = (self.pre_process_train if train else self.pre_process)((x_src, x_dst), dt)
        (x_src, x_dst), _ = self.proc((x_src, x_dst), e, edge_index, shard_info, model_comm_group if world > 1 else None)
        out = (self.post_process_train if train else self.post_process)(x_dst, dt)
        if world > 1 and not keep_x_dst_sharded:
            from ..distributed.graph import gather_rows
            from ..distributed.graph import gather_rows_grad

            out = (gather_rows_grad if train else gather_rows)(out, shard_info.dst_nodes, model_comm_group)
        return x_src, out


class GNNForwardMapper(GNNBaseMapper):
    """data -> hidden (mapper.py:863-965); returns (updated src embedding, dst)."""

    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)
        self.proc = GraphConvMapperBlock(update_src_nodes=True, **self._block_kw)
        self.emb_nodes_src = MLP(self.in_channels_src, self._mlp_hidden, self.hidden_dim, **self._mlp_kw)
        self.emb_nodes_dst = MLP(self.in_channels_dst, self._mlp_hidden, self.hidden_dim, **self._mlp_kw)

    def pre_process(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return self.emb_nodes_src.run(x[0], dt), self.emb_nodes_dst.run(x[1], dt)

    def pre_process_train(self, x: PairTensor, dt: torch.dtype) -> PairTensor:
        return T.mlp(self.emb_nodes_src, x[0], dt), T.mlp(self.emb_nodes_dst, x[1], dt)

    def forward(
        self,
        x: PairTensor,
        batch_size: int,
        shard_info: Optional[BipartiteGraphShardInfo],
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        keep_x_dst_sharded: bool = False,
        edges_are_dst_sorted: bool = True,
        **kwargs,
    ) -> PairTensor:
        return self._run(x, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted, keep_x_dst_sharded)


class GNNBackwardMapper(GNNBaseMapper):
    """hidden -> data (mapper.py:968-1087); ``pre_process`` is the identity, extractor = MLP without LayerNorm."""

    def __init__(self, **kwargs) -> None:
        super().__init__(**kwargs)
        self.proc = GraphConvMapperBlock(update_src_nodes=False, **self._block_kw)
        self.node_data_extractor = MLP(self.hidden_dim, self._mlp_hidden, self.out_channels_dst, layer_norm=False, **self._mlp_kw)

    def post_process(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return self.node_data_extractor.run(x_dst, dt)

    def post_process_train(self, x_dst: Tensor, dt: torch.dtype) -> Tensor:
        return T.mlp(self.node_data_extractor, x_dst, dt)

    def forward(
        self,
        x: PairTensor,
        batch_size: int,
        shard_info: Optional[BipartiteGraphShardInfo],
        edge_attr: Tensor,
        edge_index: Tensor,
        model_comm_group=None,
        keep_x_dst_sharded: bool = False,
        edges_are_dst_sorted: bool = True,
        **kwargs,
    ) -> Tensor:
        return self._run(x, batch_size, shard_info, edge_attr, edge_index, model_comm_group, edges_are_dst_sorted, keep_x_dst_sharded)[1]
