"""Trainable node / edge tensors and named node attributes (reference: layers/graph.py:20-118).

Same class names, constructor arguments, ``state_dict`` keys (``trainable``, ``latlons_<name>``,
``trainable_tensors.<name>.trainable``) and return values as the reference.  The reference rebuilds the
``repeat`` + ``cat`` result on every forward; here the assembled tensor is built once per parameter version and
handed back by identity afterwards, so everything keyed on the tensor identity downstream (the cached CSR plan, the
16-float padded edge rows of the attention kernel, the embedding GEMM operand) is reused from step to step.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Optional

import torch
from torch import Tensor
from torch import nn

from .._ident import version


def _repeat_rows(x: Tensor, batch_size: int) -> Tensor:
    """``einops.repeat(x, "e f -> (repeat e) f", repeat=batch_size)`` (graph.py:39-41)."""
    return x if batch_size == 1 else x.repeat(batch_size, 1)


class TrainableTensor(nn.Module):
    """``cat([x, trainable], -1)`` tiled ``batch_size`` times along the rows (graph.py:20-46)."""

    def __init__(self, tensor_size: int, trainable_size: int) -> None:
        super().__init__()
        if trainable_size > 0:
            trainable = nn.Parameter(torch.empty(tensor_size, trainable_size))
            nn.init.constant_(trainable, 0)
        else:
            trainable = None
        self.register_parameter("trainable", trainable)
        self._cache_key: Optional[tuple] = None
        self._cache_val: Optional[Tensor] = None

    def forward(self, x: Tensor, batch_size: int) -> Tensor:
        t = self.trainable
        if t is not None and torch.is_grad_enabled() and t.requires_grad:
            # someone wants gradients w.r.t. the trainable tensor: plain differentiable assembly, no cache
            return torch.cat([_repeat_rows(x, batch_size), _repeat_rows(t.to(x.device), batch_size)], dim=-1)
        key = (x.data_ptr(), version(x), tuple(x.shape), x.dtype, str(x.device), batch_size,
               None if t is None else (t.data_ptr(), version(t), str(t.device)))  # fmt: skip
        if key != self._cache_key:
            parts = [_repeat_rows(x, batch_size)]
            if t is not None:
                parts.append(_repeat_rows(t.detach().to(x.device), batch_size))
            self._cache_val = torch.cat(parts, dim=-1) if len(parts) > 1 or batch_size > 1 else x
            self._cache_key = key
        return self._cache_val


def _node_items(graph_data):
    """(name, store) pairs of the node sets of a ``HeteroData``-like object or of a plain ``{name: {"x": coords}}`` mapping."""
    if hasattr(graph_data, "node_items"):
        return list(graph_data.node_items())
    return [(k, v) for k, v in graph_data.items() if isinstance(k, str)]


def _field(store, name: str):
    return store[name] if isinstance(store, dict) and name in store else getattr(store, name)


class NamedNodesAttributes(nn.Module):
    """sin / cos of the node coordinates plus an optional trainable tensor per node set (graph.py:49-118)."""

    def __init__(self, trainable_parameters: dict[str, int], graph_data) -> None:
        super().__init__()
        trainable_parameters = defaultdict(int, trainable_parameters or {})
        items = _node_items(graph_data)
        self.num_nodes: dict[str, int] = {}
        self.attr_ndims: dict[str, int] = {}
        for name, store in items:
            x = _field(store, "x")
            n = store["num_nodes"] if isinstance(store, dict) and "num_nodes" in store else getattr(store, "num_nodes", x.shape[0])
            self.num_nodes[name] = int(n)
            self.attr_ndims[name] = 2 * x.shape[1] + trainable_parameters[name]
        self.trainable_tensors = nn.ModuleDict()
        for name, store in items:
            self.register_coordinates(name, _field(store, "x"))
            self.register_tensor(name, trainable_parameters[name])

    def register_coordinates(self, name: str, node_coords: Tensor) -> None:
        sin_cos_coords = torch.cat([torch.sin(node_coords), torch.cos(node_coords)], dim=-1)
        self.register_buffer(f"latlons_{name}", sin_cos_coords, persistent=True)

    def get_coordinates(self, name: str) -> Tensor:
        sin_cos_coords = getattr(self, f"latlons_{name}")
        ndim = sin_cos_coords.shape[1] // 2
        return torch.atan2(sin_cos_coords[:, :ndim], sin_cos_coords[:, ndim:])

    def register_tensor(self, name: str, num_trainable_params: int) -> None:
        self.trainable_tensors[name] = TrainableTensor(self.num_nodes[name], num_trainable_params)

    def forward(self, name: str, batch_size: int) -> Tensor:
        return self.trainable_tensors[name](getattr(self, f"latlons_{name}"), batch_size)
