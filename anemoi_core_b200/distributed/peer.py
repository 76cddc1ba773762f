"""Symmetric device buffers and exchange channels over NVLink peer memory (``csrc/peer.cu``, include/anemoi_b200.h).

One process per GPU on one NVSwitch box.  A symmetric buffer is a ``cudaMalloc`` allocation per rank whose CUDA-IPC handle has been sent
to every other rank of the model group (``torch.distributed`` moves the 64-byte handles once, at plan-build time); a rank then holds a
peer-mapped pointer to every other rank's copy and the exchange kernels store into them directly.  Nothing here is on the per-step host
path except three kernel launches per exchange, which a CUDA graph absorbs.

``ANEMOI_B200_PEER=0`` keeps the NCCL all-to-all (A/B switch); CPU tensors / the Gloo backend (host-logic tests) never come here.
"""

from __future__ import annotations

import atexit
import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .. import _lib
from .graph import group_rank
from .graph import group_size

ENABLED = os.environ.get("ANEMOI_B200_PEER", "1") != "0"
_OPEN: list = []  # (kind, pointer) of everything to release at exit
_STATE: dict = {}  # id(group) -> bool: can this group use peer memory?


class _DevMem:
    """``__cuda_array_interface__`` view of raw device memory (lets torch alias an IPC allocation without owning it)."""

    def __init__(self, ptr: int, nbytes: int) -> None:
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class SymmBuffer:
    """``nbytes`` of device memory on every rank of ``group``; ``ptrs[r]`` is rank r's copy as addressable from THIS process."""

    def __init__(self, nbytes: int, group) -> None:
        lib = _lib.load()
        world, me = group_size(group), group_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        nbytes = max(int(nbytes), 256)
        rc = lib.anemoi_b200_ipc_alloc(nbytes, ctypes.byref(ptr), handle)  # a failing rank still takes part in the collectives below
        if rc == 0:
            _OPEN.append(("free", ptr.value))
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
        info = torch.tensor([nbytes, 1 if rc == 0 else 0], dtype=torch.int64, device=dev)
        all_h = [torch.empty_like(mine) for _ in range(world)]
        all_n = [torch.empty_like(info) for _ in range(world)]
        dist.all_gather(all_h, mine, group=group)
        dist.all_gather(all_n, info, group=group)
        if not all(int(t[1].item()) for t in all_n):
            _lib.check(rc, "anemoi_b200_ipc_alloc")
            raise RuntimeError("anemoi_b200_ipc_alloc failed on another rank of the group")
        self.nbytes, self.sizes = nbytes, [int(t[0].item()) for t in all_n]
        self.ptrs: list[int] = []
        for r in range(world):
            if r == me:
                self.ptrs.append(ptr.value)
                continue
            h = (ctypes.c_ubyte * 64)(*all_h[r].cpu().tolist())
            p = ctypes.c_void_p()
            _lib.check(lib.anemoi_b200_ipc_open(h, ctypes.byref(p)), "anemoi_b200_ipc_open")
            _OPEN.append(("close", p.value))
            self.ptrs.append(p.value)
        self._mem = _DevMem(ptr.value, nbytes)
        self.local: Tensor = torch.as_tensor(self._mem, device=dev)  # uint8 [nbytes], aliases the allocation

    def view(self, rows: int, cols: int, dtype: torch.dtype) -> Tensor:
        n = rows * cols * torch.empty(0, dtype=dtype).element_size()
        if n > self.nbytes:
            raise ValueError("symmetric buffer too small for the requested view")
        return self.local[:n].view(dtype).view(rows, cols)


def _release() -> None:
    try:
        lib = _lib.load()
        for kind, p in reversed(_OPEN):
            (lib.anemoi_b200_ipc_close if kind == "close" else lib.anemoi_b200_ipc_free)(ctypes.c_void_p(p))
    except Exception:  # noqa: BLE001 - interpreter shutdown
        pass
    _OPEN.clear()


atexit.register(_release)


def available(group, device: Optional[torch.device] = None) -> bool:
    """Can the ranks of ``group`` exchange rows through peer memory?  Decided once per group, collectively (every rank gets the same
    answer): CUDA, NCCL backend, more than one rank, every rank able to map every other rank's test buffer."""
    if not ENABLED or group is None or group_size(group) == 1:
        return False
    key = id(group)
    if key in _STATE:
        return _STATE[key]
    ok = torch.cuda.is_available() and dist.get_backend(group) == "nccl" and (device is None or device.type == "cuda")
    if ok:
        flag = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            SymmBuffer(256, group)
        except Exception as e:  # noqa: BLE001 - no IPC between these ranks (different hosts, MIG, containers without shared /dev/shm, ...)
            import warnings

            warnings.warn(f"anemoi_core_b200: peer-memory exchange unavailable ({e}); using NCCL all-to-all")
            flag.zero_()
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        ok = bool(flag.item())
    _STATE[key] = ok
    return ok


class PeerChannel:
    """One exchange channel: the symmetric control block (exchange counters / flags) of ``csrc/peer.cu``."""

    def __init__(self, group) -> None:
        self.group, self.world, self.rank = group, group_size(group), group_rank(group)
        self.ctl = SymmBuffer(256, group)
        self.ctl_ptrs = (ctypes.c_uint64 * self.world)(*self.ctl.ptrs)

    def host_ptrs(self, values: list[int]):
        return (ctypes.c_uint64 * self.world)(*values)
