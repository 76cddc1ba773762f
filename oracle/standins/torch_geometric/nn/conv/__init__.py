"""MessagePassing.propagate semantics used by layers/conv.py (flow source_to_target)."""
import inspect

import torch
from torch import Tensor

from torch_geometric.utils import scatter


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_params = list(inspect.signature(self.message).parameters)
        self._aggr_params = list(inspect.signature(self.aggregate).parameters)

    def propagate(self, edge_index, size=None, **kwargs):
        src_idx, dst_idx = edge_index[0], edge_index[1]
        msg_kwargs = {}
        n_dst = None
        for name in self._msg_params:
            if name.endswith("_i") or name.endswith("_j"):
                base = name[:-2]
                if base == "size":  # size_i / size_j
                    continue
                data = kwargs[base]
                if isinstance(data, (tuple, list)):
                    data = data[1] if name.endswith("_i") else data[0]
                if name.endswith("_i"):
                    n_dst = data.shape[0]
                    msg_kwargs[name] = data.index_select(0, dst_idx)
                else:
                    msg_kwargs[name] = data.index_select(0, src_idx)
        if size is not None and size[1] is not None:
            n_dst = int(size[1])
        for name in self._msg_params:
            if name in msg_kwargs:
                continue
            if name == "index":
                msg_kwargs[name] = dst_idx
            elif name == "ptr":
                msg_kwargs[name] = None
            elif name == "size_i":
                msg_kwargs[name] = n_dst
            elif name in kwargs:
                msg_kwargs[name] = kwargs[name]
        out = self.message(**msg_kwargs)
        if "edge_index" in self._aggr_params:  # GraphConv overrides aggregate(edges_new, edge_index, dim_size)
            extra = {k: (edge_index if k == "edge_index" else kwargs.get(k)) for k in self._aggr_params[1:]}
            return self.aggregate(out, **extra)
        dim_size = kwargs.get("dim_size", n_dst)
        return scatter(out, dst_idx, dim=0, dim_size=dim_size, reduce="sum")

    def message(self, x_j: Tensor) -> Tensor:
        return x_j

    def aggregate(self, inputs: Tensor, index: Tensor = None, ptr=None, dim_size=None) -> Tensor:
        return scatter(inputs, index, dim=0, dim_size=dim_size, reduce="sum")
