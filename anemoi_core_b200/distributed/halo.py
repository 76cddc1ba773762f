"""Halo exchange of source rows for the dst-range sharded GraphTransformer processor.

Reference: the "edges" strategy of ``GraphTransformerProcessorBlock`` (layers/block.py:1120-1183) exchanges only the boundary rows a
rank's edges reference instead of all-gathering every source row, with the metadata cached after the first call.  Same idea here:

* every rank owns a contiguous dst / node range (balanced partition) and the edges into it (``shard_edges_1hop``); its edges name
  GLOBAL source ids;
* once per (graph, group) a ``HaloPlan`` is built: the sorted unique remote ids this rank needs from every other rank, the row lists
  every rank must send (the need-lists, exchanged once), and the edge list relabelled onto a compact table
  ``[own rows | rows from rank 0 | rows from rank 1 | ...]``;
* per layer: pack the rows to send with one gather kernel, one ``all_to_all_single`` (NCCL) straight into the tail of the table.

Traffic per layer and rank drops from (world-1)/world of the whole k|v tensor to the halo (O(10 %) on the lat-band ordered icosahedral
mesh), and the CSR plan over the compact table is built once.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .graph import SegmentedCapture
from .graph import _exchange
from .graph import group_rank
from .graph import group_size


@dataclass
class HaloPlan:
    n_local: int  # rows this rank owns
    n_halo: int  # remote rows it needs
    recv_splits: list[int]  # rows received from each rank (0 for itself)
    send_splits: list[int]  # rows sent to each rank
    send_idx: Tensor  # int32 [sum(send_splits)]: LOCAL row ids to send, grouped by destination rank
    edge_index: Tensor  # int64 [2, E_local]: src relabelled onto the compact table, dst local
    halo_ids: Tensor  # int64 [n_halo]: global id of every halo row (table row n_local + i)
    group: object
    _pending: object = None

    @property
    def n_table(self) -> int:
        return self.n_local + self.n_halo

    def exchange(self, table: Tensor) -> Tensor:
        """``table`` [n_table, W]: rows [0, n_local) hold this rank's rows; fills rows [n_local, n_table) with the halo rows."""
        if table.shape[0] != self.n_table:
            raise ValueError(f"halo exchange: table has {table.shape[0]} rows, plan needs {self.n_table}")
        local, halo = table[: self.n_local], table[self.n_local :]
        if table.is_cuda:
            from .. import ops

            send = ops.cast_pad(local, table.dtype, idx=self.send_idx)  # one gather kernel packs every destination's rows
        else:
            send = local.index_select(0, self.send_idx.long())
        if SegmentedCapture.active is not None:
            SegmentedCapture.active.exchange(send, self.send_splits, self.recv_splits, self.group, halo)
        else:
            _exchange(send, self.send_splits, self.recv_splits, self.group, out=halo)
        return table

    def exchange_start(self, table: Tensor) -> None:
        """Pack and issue the all-to-all; returns at once.  Work enqueued before ``exchange_finish`` overlaps the transfer."""
        if table.shape[0] != self.n_table:
            raise ValueError(f"halo exchange: table has {table.shape[0]} rows, plan needs {self.n_table}")
        local, halo = table[: self.n_local], table[self.n_local :]
        if table.is_cuda:
            from .. import ops

            send = ops.cast_pad(local, table.dtype, idx=self.send_idx)
        else:
            send = local.index_select(0, self.send_idx.long())
        if SegmentedCapture.active is not None:
            SegmentedCapture.active.exchange_start(send, self.send_splits, self.recv_splits, self.group, halo)
            self._pending = "capture"
        else:
            self._pending = (_exchange(send, self.send_splits, self.recv_splits, self.group, out=halo, async_op=True), send)

    def exchange_finish(self) -> None:
        pending, self._pending = self._pending, None
        if pending == "capture":
            SegmentedCapture.active.exchange_finish()
        elif pending is not None and pending[0] is not None:
            pending[0].wait()


_PLANS: dict = {}


def halo_plan_for(edge_index: Tensor, node_splits: list[int], group) -> HaloPlan:
    """Plan for the local edge list ``edge_index`` (GLOBAL src ids, LOCAL dst ids, as the processors shard it), cached on the tensor."""
    world, me = group_size(group), group_rank(group)
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), str(edge_index.device), tuple(node_splits), world, me, id(group))
    hit = _PLANS.get(key)
    if hit is not None:
        return hit[1]
    dev = edge_index.device
    bounds = [0]
    for s in node_splits:
        bounds.append(bounds[-1] + int(s))
    n_total, start, n_local = bounds[-1], bounds[me], int(node_splits[me])
    src = edge_index[0]
    uniq = torch.unique(src)  # sorted
    need, recv_splits = [], []
    for r in range(world):
        if r == me:
            recv_splits.append(0)
            continue
        ids = uniq[(uniq >= bounds[r]) & (uniq < bounds[r + 1])]
        need.append(ids)
        recv_splits.append(int(ids.numel()))
    halo_ids = torch.cat(need) if need else uniq.new_empty(0)
    # every rank tells every other rank how many and which rows it needs (once per graph and group)
    counts = torch.tensor(recv_splits, dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=group)
    send_splits = [int(all_counts[r][me]) for r in range(world)]
    wanted = _exchange(halo_ids.contiguous(), recv_splits, send_splits, group)  # global ids the others want from us, grouped by requester
    send_idx = (wanted - start).to(torch.int32).contiguous()
    if send_idx.numel() and (int(send_idx.min()) < 0 or int(send_idx.max()) >= n_local):
        raise RuntimeError("halo plan: a peer requested rows outside this rank's range (inconsistent node_splits across ranks?)")
    lut = torch.full((n_total,), -1, dtype=torch.int64, device=dev)
    lut[start : start + n_local] = torch.arange(n_local, device=dev)
    lut[halo_ids] = n_local + torch.arange(halo_ids.numel(), device=dev)
    local_ei = torch.stack([lut[src], edge_index[1]]).contiguous()
    plan = HaloPlan(n_local, int(halo_ids.numel()), recv_splits, send_splits, send_idx, local_ei, halo_ids, group)
    if len(_PLANS) > 32:
        _PLANS.clear()
    _PLANS[key] = (edge_index, plan)
    return plan
