"""Halo exchange of source rows for the dst-range sharded forward (GraphTransformer processor and mappers, GNN processor).

Reference: the "edges" strategy of ``GraphTransformerProcessorBlock`` (layers/block.py:1120-1183) exchanges only the boundary rows a
rank's edges reference instead of all-gathering every source row, with the metadata cached after the first call; the mappers all-gather
the sources and drop the unconnected ones (mapper.py:248-297, khop_edges.py:317-409).  Same idea here, for square AND bipartite graphs:

* every rank owns a contiguous range of SOURCE rows (``src_splits``) and a contiguous range of destination rows with the edges into it
  (``shard_edges_1hop``); its edges name GLOBAL source ids;
* once per (graph, group) a ``HaloPlan`` is built: the sorted unique remote ids this rank needs from every other rank, the row lists
  every rank must send (the need-lists, exchanged once), and the edge list relabelled onto a compact table
  ``[own source rows | rows from rank 0 | rows from rank 1 | ...]``;
* per exchange the producer writes this rank's rows into the head of the table (``plan.table``), then
  - CUDA + NCCL group on one box: ``csrc/peer.cu`` — rendezvous, one kernel that stores the requested rows straight into the peers'
    tables over NVLink (CUDA-IPC mapped pointers), arrival wait; three launches on the compute stream, CUDA-graph capturable;
  - otherwise (``ANEMOI_B200_PEER=0``, Gloo on CPU in the host-logic tests): one gather kernel + one ``all_to_all_single``.

Traffic per layer and rank drops from (world-1)/world of the whole k|v tensor to the halo (O(10 %) on the lat-band ordered icosahedral
mesh), and the CSR plan over the compact table is built once.
"""

from __future__ import annotations

from dataclasses import dataclass
from dataclasses import field
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from .._ident import tensor_ident  # noqa: F401  (re-exported: layers/processor.py)
from .graph import SegmentedCapture
from .graph import _exchange
from .graph import group_rank
from .graph import group_size


@dataclass
class HaloPlan:
    n_local: int  # source rows this rank owns
    n_halo: int  # remote rows it needs
    recv_splits: list[int]  # rows received from each rank (0 for itself)
    send_splits: list[int]  # rows sent to each rank
    send_idx: Tensor  # int32 [sum(send_splits)]: LOCAL row ids to send, grouped by destination rank
    edge_index: Tensor  # int64 [2, E_local]: src relabelled onto the compact table, dst local
    halo_ids: Tensor  # int64 [n_halo]: global id of every halo row (table row n_local + i)
    group: object
    peer_row0: list[int] = field(default_factory=list)  # row of rank p's table where THIS rank's rows start
    _pending: object = None
    _channel: object = None  # distributed.peer.PeerChannel (lazily, collectively)
    _send_off: Optional[Tensor] = None
    _tables: dict = field(default_factory=dict)  # (width, dtype, tag) -> (tensor, SymmBuffer | None, dst_ptrs)
    _by_ptr: dict = field(default_factory=dict)  # data_ptr of a symmetric table -> its entry

    @property
    def n_table(self) -> int:
        return self.n_local + self.n_halo

    # -- tables ----------------------------------------------------------------------------------------------------------------
    def table(self, width: int, dtype: torch.dtype, device, tag: str = "") -> Tensor:
        """The compact table [n_table, width] for this plan: a symmetric peer buffer when the group can use peer memory (allocated
        once per (width, dtype, tag), COLLECTIVELY: every rank must ask for the same tables in the same order), else a fresh tensor."""
        from . import peer

        device = torch.device(device)
        if device.type != "cuda" or not peer.available(self.group, device):
            return torch.empty((self.n_table, width), dtype=dtype, device=device)
        key = (width, dtype, tag)
        ent = self._tables.get(key)
        if ent is None:
            if self._channel is None:
                self._channel = peer.PeerChannel(self.group)
                offs = [0]
                for s in self.send_splits:
                    offs.append(offs[-1] + int(s))
                self._send_off = torch.tensor(offs, dtype=torch.int32, device=device)
            es = torch.empty(0, dtype=dtype).element_size()
            row_bytes = width * es
            if row_bytes % 16:
                raise ValueError(f"halo table rows must be multiples of 16 bytes (width {width} x {es} bytes)")
            buf = peer.SymmBuffer(max(self.n_table, 1) * row_bytes, self.group)
            t = buf.view(self.n_table, width, dtype)
            world, me = group_size(self.group), group_rank(self.group)
            dst = [0 if r == me else buf.ptrs[r] + self.peer_row0[r] * row_bytes for r in range(world)]
            ent = (t, buf, self._channel.host_ptrs(dst), row_bytes)
            self._tables[key] = ent
            self._by_ptr[t.data_ptr()] = ent
        return ent[0]

    def _peer_entry(self, table: Tensor):
        return self._by_ptr.get(table.data_ptr()) if table.is_cuda else None

    # -- exchange --------------------------------------------------------------------------------------------------------------
    def exchange(self, table: Tensor) -> Tensor:
        """``table`` [n_table, W]: rows [0, n_local) hold this rank's rows; fills rows [n_local, n_table) with the halo rows."""
        self.exchange_start(table)
        self.exchange_finish()
        return table

    def exchange_start(self, table: Tensor) -> None:
        """Send the rows the peers asked for; returns at once.  Work enqueued before ``exchange_finish`` overlaps the transfer."""
        if table.shape[0] != self.n_table:
            raise ValueError(f"halo exchange: table has {table.shape[0]} rows, plan needs {self.n_table}")
        from .. import ops

        ent = self._peer_entry(table)
        if ent is not None:
            ops.peer_rendezvous(self._channel, table)
            ops.halo_push(self._channel, table[: self.n_local], self.send_idx, self._send_off, int(self.send_idx.numel()), ent[2], ent[3])
            self._pending = ("peer", table)
            return
        local, halo = table[: self.n_local], table[self.n_local :]
        send = ops.cast_pad(local, table.dtype, idx=self.send_idx)  # one gather kernel packs every destination's rows (CUDA only, like every op)
        if SegmentedCapture.active is not None:
            SegmentedCapture.active.exchange_start(send, self.send_splits, self.recv_splits, self.group, halo)
            self._pending = ("capture", None)
        else:
            self._pending = ("nccl", (_exchange(send, self.send_splits, self.recv_splits, self.group, out=halo, async_op=True), send))

    def exchange_finish(self) -> None:
        pending, self._pending = self._pending, None
        if pending is None:
            return
        kind, payload = pending
        if kind == "peer":
            from .. import ops

            ops.halo_wait(self._channel, payload)
        elif kind == "capture":
            SegmentedCapture.active.exchange_finish()
        elif payload[0] is not None:
            payload[0].wait()


_PLANS: dict = {}


def halo_plan_for(edge_index: Tensor, src_splits: list[int], group) -> HaloPlan:
    """Plan for the local edge list ``edge_index`` (GLOBAL src ids, LOCAL dst ids, as the processors / mappers shard it) over source
    rows partitioned by ``src_splits``; cached on the tensor.  Square graphs pass the node partition."""
    world, me = group_size(group), group_rank(group)
    key = (tensor_ident(edge_index), tuple(edge_index.shape), str(edge_index.device), tuple(src_splits), world, me, id(group))
    hit = _PLANS.get(key)
    if hit is not None:
        return hit[1]
    dev = edge_index.device
    bounds = [0]
    for s in src_splits:
        bounds.append(bounds[-1] + int(s))
    n_total, start, n_local = bounds[-1], bounds[me], int(src_splits[me])
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
    all_counts = [c.tolist() for c in all_counts]
    send_splits = [int(all_counts[r][me]) for r in range(world)]
    wanted = _exchange(halo_ids.contiguous(), recv_splits, send_splits, group)  # global ids the others want from us, grouped by requester
    send_idx = (wanted - start).to(torch.int32).contiguous()
    if send_idx.numel() and (int(send_idx.min()) < 0 or int(send_idx.max()) >= n_local):
        raise RuntimeError("halo plan: a peer requested rows outside this rank's range (inconsistent src_splits across ranks?)")
    lut = torch.full((n_total,), -1, dtype=torch.int64, device=dev)
    lut[start : start + n_local] = torch.arange(n_local, device=dev)
    lut[halo_ids] = n_local + torch.arange(halo_ids.numel(), device=dev)
    local_ei = torch.stack([lut[src], edge_index[1]]).contiguous()
    # in rank p's table our rows follow p's own rows and the rows of the ranks before us
    peer_row0 = [int(src_splits[p]) + int(sum(all_counts[p][:me])) for p in range(world)]
    plan = HaloPlan(n_local, int(halo_ids.numel()), recv_splits, send_splits, send_idx, local_ei, halo_ids, group, peer_row0)
    if len(_PLANS) > 32:
        _PLANS.clear()
    _PLANS[key] = (edge_index, plan)
    return plan
