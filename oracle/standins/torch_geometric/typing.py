"""Typing aliases used by the reference hot path (torch_geometric.typing)."""
from typing import Optional, Tuple, Union

from torch import Tensor

Adj = Tensor
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
Size = Optional[Tuple[int, int]]
NoneType = Optional[Tensor]
