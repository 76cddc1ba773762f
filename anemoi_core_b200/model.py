"""Encoder -> processor -> decoder forward step (the call sequence of ``AnemoiModelEncProcDec.forward``,
models/encoder_processor_decoder.py:260-324) over the drop-in mappers / processor, plus CUDA-graph capture of the whole
step.  This is the unit the benchmark times ("forward ms/step", BASELINE.json); the surrounding model glue
(input assembly, boundings, pre/post-processors) stays with the caller.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch import nn

from . import ops
from .distributed.shapes import BipartiteGraphShardInfo
from .distributed.shapes import GraphShardInfo
from .layers import GNNBackwardMapper
from .layers import GNNForwardMapper
from .layers import GNNProcessor
from .layers import GraphTransformerBackwardMapper
from .layers import GraphTransformerForwardMapper
from .layers import GraphTransformerProcessor


class EncProcDec(nn.Module):
    """``kind`` = "graphtransformer" | "gnn".  ``graph``: dict with {enc,proc,dec}_index int64 [2,E] (dst-sorted) and
    {enc,proc,dec}_attr fp32 [E, edge_dim] (``synthetic.build_graph``)."""

    def __init__(self, kind: str, *, in_grid: int, in_mesh: int, out_grid: int, num_channels: int, num_layers: int, edge_dim: int,
                 num_heads: int = 16, mlp_hidden_ratio: float = 4.0, mlp_extra_layers: int = 0) -> None:  # fmt: skip
        super().__init__()
        self.kind = kind
        C = num_channels
        if kind == "graphtransformer":
            common = dict(num_heads=num_heads, mlp_hidden_ratio=mlp_hidden_ratio, edge_dim=edge_dim, num_chunks=1)
            self.encoder = GraphTransformerForwardMapper(in_channels_src=in_grid, in_channels_dst=in_mesh, hidden_dim=C, **common)
            self.processor = GraphTransformerProcessor(num_layers=num_layers, num_channels=C, **common)
            # the GT encoder hands back the raw grid input, which the decoder embeds (mapper.py:597, :698-701)
            self.decoder = GraphTransformerBackwardMapper(in_channels_src=C, in_channels_dst=in_grid, hidden_dim=C, out_channels_dst=out_grid, **common)
        elif kind == "gnn":
            common = dict(mlp_extra_layers=mlp_extra_layers, edge_dim=edge_dim, num_chunks=1)
            self.encoder = GNNForwardMapper(in_channels_src=in_grid, in_channels_dst=in_mesh, hidden_dim=C, **common)
            self.processor = GNNProcessor(num_layers=num_layers, num_channels=C, **common)
            # the GNN decoder's dst input is the encoder's updated src embedding (encoder_processor_decoder.py:260-269, 316-318)
            self.decoder = GNNBackwardMapper(in_channels_src=C, in_channels_dst=C, hidden_dim=C, out_channels_dst=out_grid, **common)
        else:
            raise ValueError(f"unknown model kind {kind!r}")
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    def forward(self, x_grid: Tensor, x_mesh: Tensor, graph: dict, model_comm_group=None, mesh_shards: Optional[list[int]] = None,
                grid_shards: Optional[list[int]] = None) -> Tensor:
        """One forward step.  With ``model_comm_group`` (and per-rank ``mesh_shards`` / ``grid_shards``) every stage is dst-range sharded:
        encoder (full grid sources, local mesh rows), processor (local rows, per-layer all-gather of k|v or x), latent skip, decoder (local
        mesh sources all-gathered as k|v, local grid rows); the output rows are gathered at the end.  GNN mappers run replicated."""
        from .distributed.graph import gather_rows
        from .distributed.graph import group_size
        from .distributed.graph import shard_rows

        if group_size(model_comm_group) > 1 and mesh_shards is not None:
            g = model_comm_group
            if self.kind == "graphtransformer" and grid_shards is not None:
                x_mesh_l = shard_rows(x_mesh, mesh_shards, g)
                x_grid_l = shard_rows(x_grid, grid_shards, g)
                _, x_local = self.encoder((x_grid, x_mesh_l), 1, BipartiteGraphShardInfo(src_nodes=None, dst_nodes=mesh_shards), graph["enc_attr"],
                                          graph["enc_index"], g, keep_x_dst_sharded=True)  # fmt: skip
                y_local = self.processor(x_local, 1, GraphShardInfo(nodes=mesh_shards), graph["proc_attr"], graph["proc_index"], g)
                y_local = ops.add(y_local, x_local)  # latent skip (:295-296)
                return self.decoder((y_local, x_grid_l), 1, BipartiteGraphShardInfo(src_nodes=mesh_shards, dst_nodes=grid_shards), graph["dec_attr"],
                                    graph["dec_index"], g, keep_x_dst_sharded=False)  # fmt: skip
            bi = BipartiteGraphShardInfo()
            x_data_latent, x_latent = self.encoder((x_grid, x_mesh), 1, bi, graph["enc_attr"], graph["enc_index"])
            x_local = shard_rows(x_latent, mesh_shards, g)
            y_local = self.processor(x_local, 1, GraphShardInfo(nodes=mesh_shards), graph["proc_attr"], graph["proc_index"], g)
            y_local = ops.add(y_local, x_local)
            x_proc = gather_rows(y_local, mesh_shards, g)
            return self.decoder((x_proc, x_data_latent), 1, bi, graph["dec_attr"], graph["dec_index"])
        bi = BipartiteGraphShardInfo()
        x_data_latent, x_latent = self.encoder((x_grid, x_mesh), 1, bi, graph["enc_attr"], graph["enc_index"])
        x_proc = self.processor(x_latent, 1, GraphShardInfo(nodes=[x_latent.shape[0]]), graph["proc_attr"], graph["proc_index"])
        x_proc = ops.add(x_proc, x_latent)  # latent skip (:295-296)
        return self.decoder((x_proc, x_data_latent), 1, bi, graph["dec_attr"], graph["dec_index"])

    # -- CUDA graph of one whole step ------------------------------------------------------------------------------
    def capture(self, x_grid: Tensor, x_mesh: Tensor, graph: dict, warmup: int = 2, **fwd_kwargs):
        """Capture the forward on static input buffers; returns ``replay(x_grid=None, x_mesh=None) -> Tensor`` (static output).
        ~150 kernel launches of 5-200 us each become one graph launch."""
        static_grid, static_mesh = x_grid.clone(), x_mesh.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.forward(static_grid, static_mesh, graph, **fwd_kwargs)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread polls CUDA events while we capture; in the default (global) mode that poisons the capture
        with torch.no_grad(), torch.cuda.graph(g, capture_error_mode="thread_local"):
            static_out = self.forward(static_grid, static_mesh, graph, **fwd_kwargs)
        self._graph = g

        def replay(new_grid: Optional[Tensor] = None, new_mesh: Optional[Tensor] = None) -> Tensor:
            if new_grid is not None:
                static_grid.copy_(new_grid, non_blocking=True)
            if new_mesh is not None:
                static_mesh.copy_(new_mesh, non_blocking=True)
            g.replay()
            return static_out

        return replay

    def capture_segmented(self, x_grid: Tensor, x_mesh: Tensor, graph: dict, warmup: int = 2, **fwd_kwargs):
        """Multi-GPU variant of ``capture``: the compute between two all-gathers is captured as a CUDA graph, the NCCL all-gathers run
        eagerly in between (``distributed.graph.SegmentedCapture``).  Returns ``replay(x_grid=None, x_mesh=None) -> Tensor``."""
        from .distributed.graph import SegmentedCapture

        static_grid, static_mesh = x_grid.clone(), x_mesh.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.forward(static_grid, static_mesh, graph, **fwd_kwargs)
            torch.cuda.synchronize()
            seg = SegmentedCapture()
            SegmentedCapture.active = seg
            try:
                seg.begin()
                static_out = self.forward(static_grid, static_mesh, graph, **fwd_kwargs)
                seg.end()
            finally:
                SegmentedCapture.active = None
        torch.cuda.current_stream().wait_stream(side)
        self._graph = seg

        def replay(new_grid: Optional[Tensor] = None, new_mesh: Optional[Tensor] = None) -> Tensor:
            if new_grid is not None:
                static_grid.copy_(new_grid, non_blocking=True)
            if new_mesh is not None:
                static_mesh.copy_(new_mesh, non_blocking=True)
            seg.replay()
            return static_out

        return replay
