"""Encoder -> processor -> decoder forward step (the call sequence of ``AnemoiModelEncProcDec.forward``,
models/encoder_processor_decoder.py:260-324) over the drop-in mappers / processor, plus CUDA-graph capture of the whole
step (``EncProcDec``: the unit the benchmark times, "forward ms/step", BASELINE.json), and the reference model around it
(``AnemoiModelEncProcDec``: graph providers, node attributes, input / output assembly, residual, boundings - SURVEY.md §8f ranks 1-2).
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
                grid_shards: Optional[list[int]] = None, keep_output_sharded: bool = False, inputs_sharded: bool = False) -> Tensor:
        """One forward step.  With ``model_comm_group`` (and per-rank ``mesh_shards`` / ``grid_shards``) every stage is dst-range sharded:
        encoder (full grid sources, local mesh rows), processor (local rows, per-layer all-gather of k|v or x), latent skip, decoder (local
        mesh sources all-gathered as k|v, local grid rows); the output rows are gathered at the end.  GNN mappers run replicated."""
        from .distributed.graph import gather_rows
        from .distributed.graph import gather_rows_grad
        from .distributed.graph import group_size
        from .distributed.graph import shard_rows

        def skip(a: Tensor, b: Tensor) -> Tensor:  # latent skip (:295-296); the fused add kernel carries no gradient
            return a + b.to(a.dtype) if (torch.is_grad_enabled() and (a.requires_grad or b.requires_grad)) else ops.add(a, b)

        if group_size(model_comm_group) > 1 and mesh_shards is not None:
            g = model_comm_group
            if self.kind == "graphtransformer" and grid_shards is not None:
                # ``inputs_sharded``: x_grid / x_mesh are already this rank's rows (the reference's in_out_sharded mode,
                # encoder_processor_decoder.py:176-183): nothing but the halo rows ever leaves a rank
                x_mesh_l = x_mesh if inputs_sharded else shard_rows(x_mesh, mesh_shards, g)
                x_grid_l = x_grid if inputs_sharded else shard_rows(x_grid, grid_shards, g)
                # grid rows sharded too: each rank embeds and projects (k | v) its own grid rows only, the rows its mesh nodes' edges name
                # on other ranks arrive by halo exchange (reference: shard the sources, mapper.py:248-297)
                _, x_local = self.encoder((x_grid_l, x_mesh_l), 1, BipartiteGraphShardInfo(src_nodes=grid_shards, dst_nodes=mesh_shards),
                                          graph["enc_attr"], graph["enc_index"], g, keep_x_dst_sharded=True)  # fmt: skip
                y_local = self.processor(x_local, 1, GraphShardInfo(nodes=mesh_shards), graph["proc_attr"], graph["proc_index"], g)
                y_local = skip(y_local, x_local)
                return self.decoder((y_local, x_grid_l), 1, BipartiteGraphShardInfo(src_nodes=mesh_shards, dst_nodes=grid_shards), graph["dec_attr"],
                                    graph["dec_index"], g, keep_x_dst_sharded=keep_output_sharded)  # fmt: skip
            if self.kind == "gnn" and grid_shards is not None:
                # GNN mappers sharded like the reference (mapper.py:760-835): grid and mesh rows sharded, the block all-gathers the embedded sources
                x_grid_l, x_mesh_l = shard_rows(x_grid, grid_shards, g), shard_rows(x_mesh, mesh_shards, g)
                x_data_l, x_local = self.encoder((x_grid_l, x_mesh_l), 1, BipartiteGraphShardInfo(src_nodes=grid_shards, dst_nodes=mesh_shards),
                                                 graph["enc_attr"], graph["enc_index"], g, keep_x_dst_sharded=True)  # fmt: skip
                y_local = self.processor(x_local, 1, GraphShardInfo(nodes=mesh_shards), graph["proc_attr"], graph["proc_index"], g)
                y_local = skip(y_local, x_local)
                return self.decoder((y_local, x_data_l), 1, BipartiteGraphShardInfo(src_nodes=mesh_shards, dst_nodes=grid_shards), graph["dec_attr"],
                                    graph["dec_index"], g, keep_x_dst_sharded=False)  # fmt: skip
            bi = BipartiteGraphShardInfo()
            x_data_latent, x_latent = self.encoder((x_grid, x_mesh), 1, bi, graph["enc_attr"], graph["enc_index"])
            x_local = shard_rows(x_latent, mesh_shards, g)
            y_local = self.processor(x_local, 1, GraphShardInfo(nodes=mesh_shards), graph["proc_attr"], graph["proc_index"], g)
            y_local = skip(y_local, x_local)
            x_proc = (gather_rows_grad if y_local.requires_grad else gather_rows)(y_local, mesh_shards, g)
            return self.decoder((x_proc, x_data_latent), 1, bi, graph["dec_attr"], graph["dec_index"])
        bi = BipartiteGraphShardInfo()
        x_data_latent, x_latent = self.encoder((x_grid, x_mesh), 1, bi, graph["enc_attr"], graph["enc_index"])
        x_proc = self.processor(x_latent, 1, GraphShardInfo(nodes=[x_latent.shape[0]]), graph["proc_attr"], graph["proc_index"])
        x_proc = skip(x_proc, x_latent)
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


# ------------------------------------------------------------------------------------------------------------------------
# The reference model around the step
# ------------------------------------------------------------------------------------------------------------------------
_BOUND_CODES = {"relu": 1, "leaky_relu": 2}


class AnemoiModelEncProcDec(nn.Module):
    """``anemoi.models.models.AnemoiModelEncProcDec`` (models/encoder_processor_decoder.py:33-330) with explicit constructor arguments in
    place of the hydra / DotDict configuration: the same sub-module names - ``node_attributes``, ``encoder_graph_provider[ds]``,
    ``encoder[ds]``, ``processor_graph_provider``, ``processor``, ``decoder_graph_provider[ds]``, ``decoder[ds]`` - hence the same
    ``state_dict`` keys (a reference checkpoint loads with ``strict=True``), and the same ``forward(x: dict[str, Tensor], *,
    model_comm_group=None, grid_shard_sizes=None) -> dict[str, Tensor]`` on ``(batch, time, ensemble, grid, vars)`` tensors.

    ``_assemble_input`` is one kernel that writes the embedding GEMM's operand (``ops.assemble_input``), ``_assemble_output`` one kernel
    that fuses the rearrange, the SkipConnection residual on the prognostic variables and the ReLU / LeakyReLU boundings
    (``ops.assemble_output``); the graph providers hand back identity-stable tensors so every CSR plan is built once.

    graph_data: ``{node_set: {"x": coords}}`` plus ``{(src, "to", dst): {"edge_index": ..., <attribute>: ...}}`` (a PyG ``HeteroData`` works).
    boundings: ``{dataset: [("relu" | "leaky_relu", [output variable indices]), ...]}``.
    Model sharding (both kinds): hidden rows sharded like the reference (balanced dst ranges, provider-sharded edges); the grid stays
    replicated unless ``grid_shard_sizes`` is given.
    """

    def __init__(self, kind: str, *, graph_data, dataset_names=("data",), hidden_nodes_name: str = "hidden", edge_attributes: list[str],
                 num_channels: int, n_step_input: int, n_step_output: int, num_input_channels: dict[str, int], num_output_channels: dict[str, int],
                 internal_input_idx: dict[str, list[int]], internal_output_idx: dict[str, list[int]], encoder: dict, processor: dict, decoder: dict,
                 trainable_parameters: Optional[dict[str, int]] = None, latent_skip: bool = True, boundings: Optional[dict] = None,
                 residual_step: int = -1) -> None:  # fmt: skip
        super().__init__()
        from .layers import NamedNodesAttributes
        from .layers import create_graph_provider

        if kind not in ("graphtransformer", "gnn"):
            raise ValueError(f"unknown model kind {kind!r}")
        self.kind, self.dataset_names, self._graph_name_hidden = kind, list(dataset_names), hidden_nodes_name
        self.num_channels, self.n_step_input, self.n_step_output, self.latent_skip = num_channels, n_step_input, n_step_output, latent_skip
        self.residual_step = residual_step
        tp = dict(trainable_parameters or {})
        node_sets = {k: v for k, v in graph_data.items() if isinstance(k, str)}
        self.node_attributes = NamedNodesAttributes({n: tp.get(n, 0) for n in node_sets}, node_sets)
        self._internal_input_idx = {k: list(v) for k, v in internal_input_idx.items()}
        self._internal_output_idx = {k: list(v) for k, v in internal_output_idx.items()}
        self.input_dim = {ds: n_step_input * num_input_channels[ds] + self.node_attributes.attr_ndims[ds] for ds in self.dataset_names}
        self.input_dim_latent = self.node_attributes.attr_ndims[hidden_nodes_name]
        self.output_dim = {ds: n_step_output * num_output_channels[ds] for ds in self.dataset_names}
        self.num_output_channels = dict(num_output_channels)
        n = self.node_attributes.num_nodes
        hid = hidden_nodes_name
        if kind == "graphtransformer":
            enc_cls, proc_cls, dec_cls = GraphTransformerForwardMapper, GraphTransformerProcessor, GraphTransformerBackwardMapper
        else:
            enc_cls, proc_cls, dec_cls = GNNForwardMapper, GNNProcessor, GNNBackwardMapper
        # Sub-modules are registered in the reference's order (encoder_processor_decoder.py:51-96: the three graph providers, then encoder,
        # processor, decoder): ``parameters()`` then enumerates like the reference's, which is what an optimizer state_dict indexes by, so a
        # reference training checkpoint resumes with its moments on the right tensors (tests/test_model_glue.py checks the key ORDER).
        self.encoder_graph_provider = nn.ModuleDict({
            ds: create_graph_provider(graph=graph_data[(ds, "to", hid)], edge_attributes=edge_attributes, src_size=n[ds], dst_size=n[hid],
                                      trainable_size=tp.get("data2hidden", 0)) for ds in self.dataset_names})  # fmt: skip
        self.processor_graph_provider = create_graph_provider(graph=graph_data[(hid, "to", hid)], edge_attributes=edge_attributes, src_size=n[hid],
                                                              dst_size=n[hid], trainable_size=tp.get("hidden2hidden", 0))  # fmt: skip
        self.decoder_graph_provider = nn.ModuleDict({
            ds: create_graph_provider(graph=graph_data[(hid, "to", ds)], edge_attributes=edge_attributes, src_size=n[hid], dst_size=n[ds],
                                      trainable_size=tp.get("hidden2data", 0)) for ds in self.dataset_names})  # fmt: skip
        self.encoder = nn.ModuleDict({
            ds: enc_cls(in_channels_src=self.input_dim[ds], in_channels_dst=self.input_dim_latent, hidden_dim=num_channels,
                        edge_dim=self.encoder_graph_provider[ds].edge_dim, **encoder) for ds in self.dataset_names})  # fmt: skip
        self.processor = proc_cls(num_channels=num_channels, edge_dim=self.processor_graph_provider.edge_dim, **processor)
        # GT: the decoder embeds the raw assembled grid input (mapper.py:698-701); GNN: its dst input is the encoder's src embedding
        self.decoder = nn.ModuleDict({
            ds: dec_cls(in_channels_src=num_channels, in_channels_dst=self.input_dim[ds] if kind == "graphtransformer" else num_channels,
                        hidden_dim=num_channels, out_channels_dst=self.output_dim[ds], edge_dim=self.decoder_graph_provider[ds].edge_dim, **decoder)
            for ds in self.dataset_names})  # fmt: skip
        # per-output-variable tables of the fused output kernel (plain attributes like the reference's bounding index tensors: no state_dict keys)
        self._tables: dict = {}
        self._bound_spec = {ds: list((boundings or {}).get(ds, [])) for ds in self.dataset_names}

    @classmethod
    def from_reference_config(cls, *, model_config, data_indices: dict, n_step_input: int, n_step_output: int, graph_data,
                              statistics=None) -> "AnemoiModelEncProcDec":  # fmt: skip
        """Build the model from the objects the reference's ``BaseGraphModel.__init__`` receives (models/base.py:41-97): the resolved model
        config (``model_config.model.{num_channels, model, encoder, processor, decoder, trainable_parameters, attributes, bounding,
        residual}``, attribute or key access), the per-dataset ``IndexCollection``s and the graph.  The ``_target_`` of the processor picks the
        model kind; keys the constructors of the reference take from elsewhere (``_target_``, ``trainable_size``, ``sub_graph_edge_attributes``)
        are dropped from the layer kwargs.  Only what this forward implements is accepted: ``SkipConnection`` residual, ReLU / LeakyReLU
        boundings (anything else raises)."""

        def get(obj, key, default=None):
            if isinstance(obj, dict):
                return obj.get(key, default)
            return getattr(obj, key, default)

        m = get(model_config, "model")
        proc_cfg = dict(get(m, "processor"))
        target = str(proc_cfg.get("_target_", ""))
        kind = "graphtransformer" if "GraphTransformer" in target else "gnn" if "GNN" in target else None
        if kind is None:
            raise NotImplementedError(f"processor {target!r}: only the GNN and GraphTransformer processors are part of this hot path")
        drop = ("_target_", "trainable_size", "sub_graph_edge_attributes", "_convert_", "_recursive_")

        def layer_kwargs(cfg) -> dict:
            return {k: v for k, v in dict(cfg).items() if k not in drop}

        residual = get(m, "residual")
        res_target = str(get(residual, "_target_", "SkipConnection")) if residual is not None else "SkipConnection"
        if not res_target.endswith("SkipConnection"):
            raise NotImplementedError(f"residual {res_target!r}: only SkipConnection is implemented (layers/residual.py:60-81)")
        names = list(data_indices.keys())
        boundings = {}
        for ds in names:
            name_to_index = data_indices[ds].model.output.name_to_index
            spec = []
            for b in get(m, "bounding", None) or []:
                t = str(get(b, "_target_"))
                code = "relu" if t.endswith(".ReluBounding") else "leaky_relu" if t.endswith(".LeakyReluBounding") else None
                if code is None:
                    raise NotImplementedError(f"bounding {t!r}: ReluBounding and LeakyReluBounding are implemented (layers/bounding.py:81-94)")
                wanted = set(get(b, "variables"))
                spec.append((code, [i for n, i in name_to_index.items() if n in wanted]))  # as BaseBounding._create_index (bounding.py:61-62)
            boundings[ds] = spec
        model_block = get(m, "model")
        tp = dict(get(m, "trainable_parameters") or {})
        hidden = get(model_block, "hidden_nodes_name", "hidden")
        # the reference broadcasts the "data" entry to every dataset name (base.py:78-82 broadcast_config_keys)
        tp_nodes = {**{ds: tp.get("data", 0) for ds in names}, hidden: tp.get("hidden", 0)}
        return cls(
            kind, graph_data=graph_data, dataset_names=names, hidden_nodes_name=hidden, edge_attributes=list(get(get(m, "attributes"), "edges")),
            num_channels=get(m, "num_channels"), n_step_input=n_step_input, n_step_output=n_step_output,
            num_input_channels={ds: len(data_indices[ds].model.input) for ds in names},
            num_output_channels={ds: len(data_indices[ds].model.output) for ds in names},
            internal_input_idx={ds: list(data_indices[ds].model.input.prognostic) for ds in names},
            internal_output_idx={ds: list(data_indices[ds].model.output.prognostic) for ds in names},
            encoder=layer_kwargs(get(m, "encoder")), processor=layer_kwargs(proc_cfg), decoder=layer_kwargs(get(m, "decoder")),
            trainable_parameters={**tp, **tp_nodes}, latent_skip=bool(get(model_block, "latent_skip", True)), boundings=boundings,
            residual_step=int(get(residual, "step", -1)) if residual is not None else -1,
        )  # fmt: skip

    def _output_tables(self, ds: str, device) -> tuple[Tensor, Tensor]:
        key = (ds, str(device))
        if key not in self._tables:
            v_out = self.num_output_channels[ds]
            skip = torch.full((v_out,), -1, dtype=torch.int32)
            skip[torch.tensor(self._internal_output_idx[ds], dtype=torch.long)] = torch.tensor(self._internal_input_idx[ds], dtype=torch.int32)
            bound = torch.zeros(v_out, dtype=torch.int32)
            # the reference applies its bounding layers one after the other (models/base.py, layers/bounding.py:81-94): compose them per
            # variable - relu after or before leaky_relu is relu (leaky_relu(relu(x)) = relu(leaky_relu(x)) = relu(x)), the rest is idempotent
            for name, idx in self._bound_spec[ds]:
                code = _BOUND_CODES[name]
                for i in idx:
                    bound[i] = _BOUND_CODES["relu"] if _BOUND_CODES["relu"] in (int(bound[i]), code) else code
            self._tables[key] = (skip.to(device), bound.to(device))
        return self._tables[key]

    def _assemble_input(self, x: Tensor, batch_size: int, grid_shard_sizes, model_comm_group, dataset_name: str, dt: torch.dtype,
                        differentiable: bool = False):  # fmt: skip
        from .distributed.graph import shard_rows
        from .layers._functional import pad_k

        attrs = self.node_attributes(dataset_name, batch_size=batch_size)
        sizes = grid_shard_sizes[dataset_name] if grid_shard_sizes is not None else None
        if sizes is not None:
            attrs = shard_rows(attrs, sizes, model_comm_group)
        if differentiable:
            # training: the reference's own statement (encoder_processor_decoder.py:115-125) in PyTorch, so that autograd reaches the input and the
            # trainable node tensor; the mappers' differentiable path casts / pads its GEMM operands itself
            b, t, e, g, v = x.shape
            rows = x.permute(0, 2, 3, 1, 4).reshape(b * e * g, t * v).float()
            return torch.cat([rows, attrs.float().repeat(rows.shape[0] // max(attrs.shape[0], 1), 1)], dim=-1), sizes
        k = x.shape[1] * x.shape[4] + attrs.shape[1]
        return ops.assemble_input(x, attrs, dt, k_pad=pad_k(k, dt)), sizes

    def _assemble_output_differentiable(self, dec: Tensor, x: Tensor, batch_size: int, ensemble_size: int, dataset_name: str) -> Tensor:
        """Training: ``_assemble_output`` (encoder_processor_decoder.py:129-163: rearrange, SkipConnection residual of the input's ``residual_step``
        slice on the prognostic variables, the bounding layers) as out-of-place PyTorch operations on the same per-variable tables the fused
        inference kernel reads."""
        skip, bound = self._output_tables(dataset_name, dec.device)
        g = dec.shape[0] // (batch_size * ensemble_size)
        y = dec.float().reshape(batch_size, ensemble_size, g, self.n_step_output, -1).permute(0, 3, 1, 2, 4)
        has = skip >= 0
        res = x[:, self.residual_step].float().unsqueeze(1)[..., skip.clamp_min(0).long()] * has.to(y.dtype)  # [B, 1, E, G, V_out], zero where no residual
        y = y + res
        y = torch.where(bound == 1, torch.relu(y), torch.where(bound == 2, torch.nn.functional.leaky_relu(y), y))
        return y.to(x.dtype)

    def forward(self, x: dict[str, Tensor], *, model_comm_group=None, grid_shard_sizes=None, **kwargs) -> dict[str, Tensor]:
        from .distributed.balanced_partition import get_balanced_partition_sizes
        from .distributed.graph import group_size
        from .distributed.graph import shard_rows
        from .layers._functional import compute_dtype
        from .layers._train import wants_grad

        names = list(x.keys())
        batch = {t.shape[0] for t in x.values()}
        ens = {t.shape[2] for t in x.values()}
        assert len(batch) == 1 and len(ens) == 1, "Dimensions must be the same across datasets"
        batch_size, ensemble_size = batch.pop(), ens.pop()
        world = group_size(model_comm_group)
        if world > 1:
            assert batch_size == 1 and ensemble_size == 1, "Only batch / ensemble size 1 per device when the model is sharded across GPUs"
        dt = compute_dtype(*x.values())
        # differentiable glue (PyTorch statements of the two assembly kernels and of the adds) whenever gradients are wanted: the fused kernels
        # have no backward.  The mappers / processor switch to their own differentiable paths by the same rule.
        train = wants_grad(self, *x.values())
        if train and world > 1:
            raise NotImplementedError("training AnemoiModelEncProcDec on a model-parallel group is not implemented: use the reference model class with the "
                                      "B200 layer classes (INTEGRATION.md §1), whose processors and mappers train sharded")  # fmt: skip
        add = (lambda a, b: a + b.to(a.dtype)) if train else ops.add
        hid = self._graph_name_hidden
        x_hidden = self.node_attributes(hid, batch_size=batch_size)
        sizes_hidden = get_balanced_partition_sizes(x_hidden.shape[0], world) if world > 1 else None
        x_hidden = shard_rows(x_hidden, sizes_hidden, model_comm_group)
        latents, x_data_latents, sizes_data = [], {}, {}
        for ds in names:
            x_data, sizes_data[ds] = self._assemble_input(x[ds], batch_size, grid_shard_sizes, model_comm_group, ds, dt, differentiable=train)
            ea, ei, es = self.encoder_graph_provider[ds].get_edges(batch_size=batch_size, model_comm_group=model_comm_group,
                                                                   shard_edges=getattr(self.encoder[ds], "shard_strategy", "edges") != "heads")  # fmt: skip
            info = BipartiteGraphShardInfo(src_nodes=sizes_data[ds], dst_nodes=sizes_hidden, edges=es)
            x_data_latents[ds], lat = self.encoder[ds]((x_data, x_hidden), batch_size, info, ea, ei, model_comm_group, keep_x_dst_sharded=True)
            latents.append(lat)
        x_latent = latents[0]
        for lat in latents[1:]:
            x_latent = add(x_latent, lat)
        # the heads (Ulysses) strategy attends over the FULL edge list for its heads: ask the provider not to cut it
        ea, ei, es = self.processor_graph_provider.get_edges(batch_size=batch_size, model_comm_group=model_comm_group,
                                                             shard_edges=getattr(self.processor, "shard_strategy", "edges") != "heads")  # fmt: skip
        x_proc = self.processor(x_latent, batch_size, GraphShardInfo(nodes=sizes_hidden if world > 1 else [x_latent.shape[0]], edges=es), ea, ei,
                                model_comm_group)  # fmt: skip
        if self.latent_skip:
            x_proc = add(x_proc, x_latent)
        out = {}
        for ds in names:
            ea, ei, es = self.decoder_graph_provider[ds].get_edges(batch_size=batch_size, model_comm_group=model_comm_group,
                                                                   shard_edges=getattr(self.decoder[ds], "shard_strategy", "edges") != "heads")  # fmt: skip
            info = BipartiteGraphShardInfo(src_nodes=sizes_hidden, dst_nodes=sizes_data[ds], edges=es)
            if world > 1 and sizes_data[ds] is None:
                # replicated grid: every rank owns a balanced slice of the grid rows inside the decoder and the output is gathered
                n_grid = x[ds].shape[3]
                info = BipartiteGraphShardInfo(src_nodes=sizes_hidden, dst_nodes=get_balanced_partition_sizes(n_grid, world), edges=None)
                ea, ei, _ = self.decoder_graph_provider[ds].get_edges(batch_size=batch_size, model_comm_group=model_comm_group, shard_edges=False)
                x_dst = x_data_latents[ds]
                if x_dst.shape[0] == n_grid:  # GraphTransformer encoder: the raw assembled input came back whole; the GNN encoder cut its
                    x_dst = shard_rows(x_dst, info.dst_nodes, model_comm_group)  # replicated source to the same balanced slice itself
                dec = self.decoder[ds]((x_proc, x_dst), batch_size, info, ea, ei, model_comm_group, keep_x_dst_sharded=False)
            else:
                dec = self.decoder[ds]((x_proc, x_data_latents[ds]), batch_size, info, ea, ei, model_comm_group,
                                       keep_x_dst_sharded=sizes_data[ds] is not None)  # fmt: skip
            if train:
                out[ds] = self._assemble_output_differentiable(dec, x[ds], batch_size, ensemble_size, ds)
                continue
            skip, bound = self._output_tables(ds, dec.device)
            out[ds] = ops.assemble_output(dec, x[ds], batch_size, ensemble_size, self.n_step_output, self.residual_step, skip, bound).to(x[ds].dtype)
        return out
