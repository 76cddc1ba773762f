import torch
from torch import Tensor


def index2ptr(index: Tensor, size=None) -> Tensor:
    if size is None:
        size = int(index.max()) + 1 if index.numel() > 0 else 0
    return torch._convert_indices_from_coo_to_csr(index, int(size), out_int32=index.dtype != torch.int64)
