"""Cache identities of tensors (per-graph plans, derived weights, row-statistics tags are cached on the identity of their source tensors).

``version(t)`` is the in-place version counter, except for INFERENCE tensors (created under ``torch.inference_mode()``, which Lightning's
validate / test / predict loops enter by default): those track no version and ``t._version`` raises — they report -1 (nothing mutates a graph or
an activation in place between our producer and consumer, and the cache entries hold the tensors, so the storage pointer still identifies them).
``tensor_ident(t)`` is (storage pointer, version) — or, for an EMPTY tensor, the object itself: every empty tensor has the same data_ptr, so a rank
whose rows receive no edge would share the cached split / plan of an earlier graph with the same partition while its peers miss and enter the
plan's collectives alone (a hang; tests/test_sharded_forward_gloo.py::test_degenerate_graphs).  Entries hold the tensor, so the id is not recycled."""
from torch import Tensor


def version(t: Tensor) -> int:
    return -1 if t.is_inference() else t._version


def tensor_ident(t: Tensor) -> tuple:
    return (t.data_ptr(), version(t)) if t.numel() else ("empty", id(t))
