"""PyG >= 2.3 utils semantics restated for the oracle."""
from typing import Optional

import torch
from torch import Tensor


def _bcast(index: Tensor, src: Tensor, dim: int) -> Tensor:
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.shape)
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _bcast(index, src, dim), src)
    if reduce == "max":
        return src.new_zeros(size).scatter_reduce_(dim, _bcast(index, src, dim), src, "amax", include_self=False)
    raise NotImplementedError(reduce)


def softmax(src: Tensor, index: Optional[Tensor] = None, ptr: Optional[Tensor] = None, num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
    assert index is not None and ptr is None
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    src_max = scatter(src.detach(), index, dim, dim_size=N, reduce="max")
    out = src - src_max.index_select(dim, index)
    out = out.exp()
    out_sum = scatter(out, index, dim, dim_size=N, reduce="sum") + 1e-16
    out_sum = out_sum.index_select(dim, index)
    return out / out_sum


def index_sort(inputs: Tensor, max_value: Optional[int] = None, stable: bool = False):
    return inputs.sort(stable=stable)


def degree(index: Tensor, num_nodes: Optional[int] = None, dtype=None) -> Tensor:
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros((N,), dtype=dtype, device=index.device)
    one = torch.ones((index.size(0),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index, one)


def mask_to_index(mask: Tensor) -> Tensor:
    return mask.nonzero(as_tuple=False).view(-1)


def bipartite_subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, size=None, return_edge_mask=False):
    """Only the non-relabelling form used by khop_edges.py's slow path."""
    assert not relabel_nodes
    src_subset, dst_subset = subset
    n_src, n_dst = size if size is not None else (int(edge_index[0].max()) + 1, int(edge_index[1].max()) + 1)

    def as_mask(s, n):
        if s.dtype == torch.bool:
            return s
        m = torch.zeros(n, dtype=torch.bool, device=edge_index.device)
        m[s] = True
        return m

    edge_mask = as_mask(src_subset, n_src)[edge_index[0]] & as_mask(dst_subset, n_dst)[edge_index[1]]
    ei = edge_index[:, edge_mask]
    ea = edge_attr[edge_mask] if edge_attr is not None else None
    if return_edge_mask:
        return ei, ea, edge_mask
    return ei, ea


def k_hop_subgraph(node_idx, num_hops, edge_index, relabel_nodes=False, num_nodes=None, flow="source_to_target", directed=False):
    """1-hop, non-relabelling form (khop_edges.py slow path)."""
    assert num_hops == 1 and not relabel_nodes
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    row, col = (edge_index[0], edge_index[1]) if flow == "source_to_target" else (edge_index[1], edge_index[0])
    node_mask = torch.zeros(num_nodes, dtype=torch.bool, device=edge_index.device)
    node_idx = torch.as_tensor(node_idx, device=edge_index.device).view(-1)
    node_mask[node_idx] = True
    edge_mask = node_mask[col]
    subset = torch.cat([node_idx, row[edge_mask]]).unique()
    if not directed:
        nm = torch.zeros(num_nodes, dtype=torch.bool, device=edge_index.device)
        nm[subset] = True
        edge_mask = nm[row] & nm[col]
    return subset, edge_index[:, edge_mask], None, edge_mask
