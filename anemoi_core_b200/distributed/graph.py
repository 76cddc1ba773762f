"""Forward halves of the reference's sharding collectives (distributed/graph.py:66-253, primitives.py:144-183) for the
dst-range ("edges") strategy: every rank owns a contiguous range of destination rows and all edges into them
(khop_edges.py:266-314), so no cross-rank reduction is needed — only source rows move.

One process per GPU, ``torch.distributed`` process group (NCCL over NVLink/NVSwitch on the B200 box, Gloo in the
CPU tests).  With ``model_comm_group`` None or of size 1 every function is the identity.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor


def group_size(group: Optional[dist.ProcessGroup]) -> int:
    return 1 if group is None else dist.get_world_size(group=group)


def group_rank(group: Optional[dist.ProcessGroup]) -> int:
    return 0 if group is None else dist.get_rank(group=group)


def shard_rows(x: Tensor, sizes: Optional[list[int]], group: Optional[dist.ProcessGroup]) -> Tensor:
    """Local row range of a replicated tensor (reference ``shard_tensor`` forward, graph.py:66-91)."""
    if group_size(group) == 1 or sizes is None:
        return x
    r = group_rank(group)
    start = sum(sizes[:r])
    return x[start : start + sizes[r]]


def _exchange(send: Tensor, in_splits: list[int], out_splits: list[int], group, out: Optional[Tensor] = None, async_op: bool = False):
    """Variable-size all-to-all of row blocks: rank r receives ``out_splits[s]`` rows from every rank s (``in_splits`` = rows sent to each).
    Returns ``out``; with ``async_op`` (NCCL) the work handle instead - the collective runs on NCCL's stream and the caller's stream only
    joins it at ``work.wait()``, so kernels enqueued in between overlap the transfer."""
    if out is None:
        out = send.new_empty((sum(out_splits),) + tuple(send.shape[1:]))
    if dist.get_backend(group) == "nccl":
        work = dist.all_to_all_single(out, send, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group, async_op=async_op)
        return work if async_op else out
    # Gloo has no all_to_all: pairwise isend / irecv (host-logic tests on CPU)
    world, me = group_size(group), group_rank(group)
    reqs, so, ro = [], 0, 0
    send_parts, recv_parts = [], []
    for r in range(world):
        send_parts.append(send[so : so + in_splits[r]])
        recv_parts.append(out[ro : ro + out_splits[r]])
        so += in_splits[r]
        ro += out_splits[r]
    recv_parts[me].copy_(send_parts[me])
    for r in range(world):
        if r == me:
            continue
        peer = dist.get_global_rank(group, r) if group is not None else r
        if in_splits[r]:
            reqs.append(dist.isend(send_parts[r].contiguous(), peer, group=group))
        if out_splits[r]:
            reqs.append(dist.irecv(recv_parts[r], peer, group=group))
    for q in reqs:
        q.wait()
    return None if async_op else out


class SegmentedCapture:
    """CUDA-graph capture of a sharded forward as a chain [graph_0, all-gather, graph_1, all-gather, ...]: the compute between two
    collectives is captured (one graph launch instead of dozens of kernel launches), the NCCL all-gathers run eagerly between the
    replays.  (Capturing the NCCL calls themselves hung on the 2-GPU box in round 1.)  ``gather_rows`` consults the active instance."""

    active: "Optional[SegmentedCapture]" = None

    def __init__(self) -> None:
        self.items: list = []
        self.pool = torch.cuda.graph_pool_handle()
        self._g: Optional[torch.cuda.CUDAGraph] = None
        self._work = None

    def begin(self) -> None:
        self._g = torch.cuda.CUDAGraph()
        self._g.capture_begin(pool=self.pool, capture_error_mode="thread_local")

    def end(self) -> None:
        self._g.capture_end()
        self.items.append(("graph", self._g))
        self._g = None

    def gather(self, x: Tensor, sizes: list[int], group) -> Tensor:
        self.end()
        out = _all_gather_rows(x, sizes, group)  # eager: allocates the persistent output and exercises the collective
        self.items.append(("gather", out, x, list(sizes), group))
        self.begin()
        return out

    def exchange(self, send: Tensor, in_splits: list[int], out_splits: list[int], group, out: Tensor) -> Tensor:
        """Halo all-to-all between two graph segments (``send`` and ``out`` are static buffers of the capture pool / the caller)."""
        self.end()
        _exchange(send, in_splits, out_splits, group, out=out)
        self.items.append(("exchange", out, send, list(in_splits), list(out_splits), group))
        self.begin()
        return out

    def exchange_start(self, send: Tensor, in_splits: list[int], out_splits: list[int], group, out: Tensor) -> None:
        """Split form of ``exchange``: the all-to-all is issued here and joined at ``exchange_finish``; the graph segment captured in
        between (the q | self GEMM of the block) runs while the rows travel."""
        self.end()
        self._work = _exchange(send, in_splits, out_splits, group, out=out, async_op=True)
        self.items.append(("xstart", out, send, list(in_splits), list(out_splits), group))
        self.begin()

    def exchange_finish(self) -> None:
        self.end()
        if self._work is not None:
            self._work.wait()
        self.items.append(("xwait",))
        self.begin()

    def replay(self) -> None:
        work = None
        for it in self.items:
            if it[0] == "graph":
                it[1].replay()
            elif it[0] == "xstart":
                _, out, send, in_splits, out_splits, group = it
                work = _exchange(send, in_splits, out_splits, group, out=out, async_op=True)
            elif it[0] == "xwait":
                if work is not None:
                    work.wait()
                    work = None
            elif it[0] == "exchange":
                _, out, send, in_splits, out_splits, group = it
                _exchange(send, in_splits, out_splits, group, out=out)
            else:
                _, out, x, sizes, group = it
                _all_gather_rows(x, sizes, group, out=out)


def gather_rows(x: Tensor, sizes: Optional[list[int]], group: Optional[dist.ProcessGroup]) -> Tensor:
    """All-gather row shards into the full tensor (reference ``gather_tensor`` / ``sync_tensor`` forward,
    graph.py:94-135, primitives.py:144-183).  Shards may differ by one row (balanced partition); NCCL's
    all_gather_into_tensor needs equal sizes, so unequal shards go through the list form."""
    world = group_size(group)
    if world == 1:
        return x
    if sizes is None:
        raise ValueError("gather_rows: per-rank shard sizes are required when the model group has more than one rank")
    if len(sizes) != world or x.shape[0] != sizes[group_rank(group)]:
        raise ValueError(f"gather_rows: local shard has {x.shape[0]} rows, shard sizes {sizes} (world {world})")
    if SegmentedCapture.active is not None:
        return SegmentedCapture.active.gather(x.contiguous(), sizes, group)
    return _all_gather_rows(x.contiguous(), sizes, group)


def _all_gather_rows(x: Tensor, sizes: list[int], group, out: Optional[Tensor] = None) -> Tensor:
    world = len(sizes)
    if out is None:
        out = torch.empty((sum(sizes),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, x, group=group)
    elif dist.get_backend(group) == "nccl":
        dist.all_gather(list(torch.split(out, sizes, dim=0)), x, group=group)  # NCCL handles unequal shards (grouped send/recv)
    else:
        # Gloo needs equal shapes: pad every shard to the largest one, like the reference (primitives.py:170-183)
        m = max(sizes)
        padded = x.new_zeros((m,) + tuple(x.shape[1:]))
        padded[: x.shape[0]] = x
        buf = x.new_empty((world * m,) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(buf, padded, group=group)
        off = 0
        for r, n in enumerate(sizes):
            out[off : off + n] = buf[r * m : r * m + n]
            off += n
    return out


# ------------------------------------------------------------------------------------------------------------------------------------
# autograd halves (reference: distributed/graph.py:227-500 ``_SyncParallelSection`` & co., block.py:1159-1169 halo exchange) — the training
# path of a model-parallel group.  Forward = the collective above; backward = its transpose.
# ------------------------------------------------------------------------------------------------------------------------------------
class GatherRowsFn(torch.autograd.Function):
    """All-gather of row shards; backward: every rank sums the cotangents of ITS rows over the group (all-reduce, keep own slice)."""

    @staticmethod
    def forward(ctx, x: Tensor, sizes: list, group) -> Tensor:
        ctx.sizes, ctx.group = list(sizes), group
        return _all_gather_rows(x.contiguous(), list(sizes), group)

    @staticmethod
    def backward(ctx, d_full: Tensor):
        d_full = d_full.contiguous()
        dist.all_reduce(d_full, group=ctx.group)
        r = group_rank(ctx.group)
        start = sum(ctx.sizes[:r])
        return d_full[start : start + ctx.sizes[r]], None, None


class HaloExchangeFn(torch.autograd.Function):
    """Halo rows of a ``HaloPlan`` from this rank's rows; backward: the halo cotangents travel back to the rows' owners and are added there
    (per requesting rank in rank order: deterministic; a requester's list holds every row once)."""

    @staticmethod
    def forward(ctx, local: Tensor, plan) -> Tensor:
        from .. import ops

        ctx.plan, ctx.n_local = plan, local.shape[0]
        send = ops.cast_pad(local.detach(), local.dtype, idx=plan.send_idx)
        return _exchange(send, plan.send_splits, plan.recv_splits, plan.group)

    @staticmethod
    def backward(ctx, d_halo: Tensor):
        plan = ctx.plan
        back = _exchange(d_halo.contiguous(), plan.recv_splits, plan.send_splits, plan.group)  # cotangents of the rows WE sent, grouped by requester
        d_local = back.new_zeros((ctx.n_local, back.shape[1]))
        off = 0
        idx = plan.send_idx.long()
        for n in plan.send_splits:
            if n:
                d_local[idx[off : off + n]] += back[off : off + n]
            off += n
        return d_local, None


class AllToAllFn(torch.autograd.Function):
    """Variable-size all-to-all of row blocks (``_exchange``); backward: the same exchange with the split lists swapped.  The differentiable
    half of the heads (Ulysses) strategy's two exchanges (reference distributed/graph.py: ``shard_heads`` / ``shard_sequence`` pairs)."""

    @staticmethod
    def forward(ctx, send: Tensor, in_splits: list, out_splits: list, group) -> Tensor:
        ctx.meta = (list(in_splits), list(out_splits), group)
        return _exchange(send.detach().contiguous(), list(in_splits), list(out_splits), group)

    @staticmethod
    def backward(ctx, d_out: Tensor):
        in_splits, out_splits, group = ctx.meta
        return _exchange(d_out.contiguous(), out_splits, in_splits, group), None, None, None


def gather_rows_grad(x: Tensor, sizes: Optional[list[int]], group) -> Tensor:
    """Differentiable ``gather_rows``."""
    if group_size(group) == 1:
        return x
    return GatherRowsFn.apply(x, sizes, group)
