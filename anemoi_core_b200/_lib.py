"""ctypes binding of libanemoi_b200.so (the C ABI declared in include/anemoi_b200.h).

The library is the product: there is no Python / PyTorch / CPU fallback.  If the shared object is missing or a
call fails, a RuntimeError is raised with the library's own message.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p
from ctypes import c_float
from ctypes import c_int
from ctypes import c_int64
from ctypes import c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# ANEMOI_B200_LIB selects another build of the same library (A/B kernel experiments); the default is the in-tree build.
LIB_PATH = os.environ.get("ANEMOI_B200_LIB") or os.path.join(_HERE, "lib", "libanemoi_b200.so")

F32, BF16 = 0, 1
EPI_GELU = 1
EPI_REVERSE = 2

# name -> argtypes (restype is int unless stated).  Must list every symbol of include/anemoi_b200.h
# (tests/test_abi.py parses the header and checks both directions).
SIGNATURES = {
    "anemoi_b200_abi_version": [],
    "anemoi_b200_last_error": [],
    "anemoi_b200_csr_build": [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "anemoi_b200_layer_norm": [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64,
                               c_int64, c_int64, c_float, c_void_p],
    "anemoi_b200_cond_layer_norm": [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                    c_int64, c_int64, c_int64, c_float, c_void_p],
    "anemoi_b200_linear": [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                           c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p],
    "anemoi_b200_row_stats": [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_float, c_int, c_void_p],
    "anemoi_b200_gt_attention_fwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p],
    "anemoi_b200_gt_attention_bwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p],
    "anemoi_b200_layer_norm_bwd": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_void_p],
    "anemoi_b200_segment_sum": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p],
    "anemoi_b200_col_sum": [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p],
    "anemoi_b200_gelu": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p],
    "anemoi_b200_attn_tile_plan": [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "anemoi_b200_gt_attention_tiled_fwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                           c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                           c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p],
    "anemoi_b200_ipc_alloc": [c_int64, c_void_p, c_void_p],
    "anemoi_b200_ipc_open": [c_void_p, c_void_p],
    "anemoi_b200_ipc_close": [c_void_p],
    "anemoi_b200_ipc_free": [c_void_p],
    "anemoi_b200_peer_rendezvous": [c_void_p, c_int64, c_int64, c_void_p],
    "anemoi_b200_halo_push": [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p],
    "anemoi_b200_halo_wait": [c_void_p, c_int64, c_int64, c_void_p],
    "anemoi_b200_graphconv_ln_aggregate": [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                           c_int64, c_int64, c_int64, c_float, c_int, c_void_p],
    "anemoi_b200_graphconv_fused": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int,
                                    c_void_p],
    "anemoi_b200_cast_pad": [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64, c_void_p],
    "anemoi_b200_glu_combine": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p],
    "anemoi_b200_glu_combine_bwd": [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p],
    "anemoi_b200_split_bf16x3": [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p],
    "anemoi_b200_assemble_input": [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64,
                                   c_int, c_void_p],
    "anemoi_b200_assemble_output": [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_int64, c_void_p],
    "anemoi_b200_add": [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_void_p],
}  # fmt: skip

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library has not been built (run __graft_entry__.build() or "
                "`make -C anemoi_core_b200/csrc`). There is no CPU or PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = c_char_p if name == "anemoi_b200_last_error" else c_int
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().anemoi_b200_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")
