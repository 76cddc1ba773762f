"""Processor / mapper blocks (reference: layers/block.py:275-479 GraphConv*, :482-1273 GraphTransformer*).

Same constructor kwargs, forward signatures, return values and ``state_dict`` keys as the reference blocks.
Forward = a short, fixed sequence of fused sm_100a kernels (DESIGN.md lists them per block); single-GPU or
dst-range-sharded across a model communication group (``distributed/graph.py``).
"""

from __future__ import annotations

import os
from typing import Optional
from typing import Union

import torch
from torch import Tensor
from torch import nn

from .. import ops
from ..distributed.graph import gather_rows
from ..distributed.graph import gather_rows_grad
from ..distributed.graph import group_size
from ..distributed.halo import halo_plan_for
from ..distributed.shapes import BipartiteGraphShardInfo
from ..distributed.shapes import GraphShardInfo
from . import _functional as Fn
from . import _train as T
from .conv import GraphConv
from .mlp import MLP
from .utils import compute_mlp_hidden_dim
from .utils import load_layer_kernels

PairTensor = tuple[Tensor, Tensor]
HALO_EXCHANGE = os.environ.get("ANEMOI_B200_HALO", "1") != "0"  # sharded GraphTransformer processor: halo all-to-all (default) vs all-gather
# Destination-tile tensor-core attention kernel (csrc/attention_tile.cu): "0" (default) = warp-per-node kernel only, "auto" = when the
# graph's tile plan re-uses gathered source rows (locality-ordered processor graphs: 2.65 on the Hilbert-ordered ico-6 mesh), "1" = whenever
# a plan exists.  Opt-in because it is parity-green but measured SLOWER at cfg2 (221 us vs 130 us, profiles/r2/README.md: 2 CTAs of 8
# warps per SM cannot hide its three dependent round trips, and the per-edge attribute phases keep it at 1 790 instructions per head-tile).
ATTN_TILES = os.environ.get("ANEMOI_B200_ATTN_TILES", "0")
ATTN_TILES_MIN_REUSE = float(os.environ.get("ANEMOI_B200_ATTN_TILES_MIN_REUSE", "1.3"))


def attention_tile_plan(csr, channels: int, heads: int, dt: torch.dtype, dp: int):
    """The tile plan to run attention with, or None for the warp-per-node kernel."""
    if ATTN_TILES == "0" or not ops.attention_tiles_supported(channels, heads, dt, dp):
        return None
    plan = ops.attention_tiles(csr)
    if plan is None or (ATTN_TILES != "1" and plan.reuse < ATTN_TILES_MIN_REUSE):
        return None
    return plan


# ------------------------------------------------------------------------------------------------------------
# GraphConv (GNN) blocks
# ------------------------------------------------------------------------------------------------------------
# L2-aware traversal order of the GraphTransformer block's kernels (ops.set_traversal): qkv GEMM top-down, attention bottom-up, projection
# top-down, row statistics bottom-up, MLP-1 top-down, MLP-2 bottom-up, next block's row statistics top-down ... each kernel starts on the
# rows its producer wrote last.  ANEMOI_B200_SERPENTINE=0 restores top-down everywhere (A/B).
SERPENTINE = os.environ.get("ANEMOI_B200_SERPENTINE", "1") != "0"

class GraphConvBaseBlock(nn.Module):
    """Edge-MLP message passing + node MLP (block.py:275-358)."""

    def __init__(
        self,
        *,
        in_channels: int,
        out_channels: int,
        num_chunks: int = 1,
        mlp_extra_layers: int = 0,
        mlp_hidden_ratio: float = 1.0,
        mlp_implementation: str = "mlp",
        update_src_nodes: bool = True,
        layer_kernels=None,
        edge_dim: Optional[int] = None,
        **kwargs,
    ) -> None:
        super().__init__()
        layer_kernels = load_layer_kernels(layer_kernels)
        hidden_dim = compute_mlp_hidden_dim(out_channels, mlp_hidden_ratio)
        mlp_kw = dict(layer_kernels=layer_kernels, n_extra_layers=mlp_extra_layers + 1, mlp_implementation=mlp_implementation)
        self.emb_edges = MLP(edge_dim, hidden_dim, out_channels, **mlp_kw) if edge_dim else None
        self.update_src_nodes = update_src_nodes
        self.num_chunks = num_chunks  # edge chunking only bounds the reference's [E, 3C] temporaries; nothing to bound here
        self.node_mlp = MLP(2 * in_channels, hidden_dim, out_channels, **mlp_kw)
        self.conv = GraphConv(in_channels=in_channels, out_channels=out_channels, layer_kernels=layer_kernels, mlp_extra_layers=mlp_extra_layers,
                              mlp_implementation=mlp_implementation)  # fmt: skip
        self.in_channels, self.out_channels = in_channels, out_channels

    def _node_update(self, x: Tensor, agg_buf: Tensor, dt: torch.dtype) -> Tensor:
        """node_mlp(cat[x, out]) + x, with ``out`` already sitting in the right half of ``agg_buf`` [N, 2C]."""
        C = self.in_channels
        ops.cast_pad(x, dt, out=agg_buf[:, :C])
        return self.node_mlp.run(agg_buf, dt, residual=agg_buf[:, :C])


class GraphConvProcessorBlock(GraphConvBaseBlock):
    def forward(
        self,
        x: Tensor,
        edge_attr: Tensor,
        edge_index: Tensor,
        shard_info: Optional[GraphShardInfo] = None,
        model_comm_group=None,
        size=None,
        **layer_kwargs,
    ) -> tuple[Tensor, Tensor]:
        dt = Fn.compute_dtype(x, edge_attr)
        if T.wants_grad(self, x, edge_attr):  # differentiable path (layers/_train.py)
            # model-parallel: all source rows by a differentiable all-gather (block.py:375); the processor hands us edges with LOCAL dst ids
            x_all = gather_rows_grad(x, shard_info.nodes if shard_info is not None else None, model_comm_group)
            (_, x_new), edges_new = T.gnn_block(self, x_all, x, edge_attr, edge_index, dt, bipartite=False)
            return x_new, edges_new
        if self.emb_edges is not None:
            edge_attr = self.emb_edges.run(edge_attr, dt)
        halo_plan = layer_kwargs.get("halo_plan")
        if halo_plan is not None:
            # halo form (distributed/halo.py): own rows + the remote source rows our edges name in one compact table; the edge list was
            # relabelled onto it (local dst ids), so the node-level projections run over n_local + n_halo rows instead of all N and the
            # aggregate is local from the start.  ``edge_index`` is ignored in favour of the plan's.
            C, n_local = self.in_channels, x.shape[0]
            table = halo_plan.table(C, dt, x.device, tag="gnn")
            ops.cast_pad(x, dt, out=table[:n_local])
            halo_plan.exchange(table)
            csr = Fn.csr_for(halo_plan.edge_index, halo_plan.n_table, n_local)
            agg_buf = torch.empty((n_local, 2 * C), dtype=dt, device=x.device)
            _, edges_new = self.conv.run(table, table[:n_local], edge_attr, csr, dt, out=agg_buf[:, C:])
            return self._node_update(x, agg_buf, dt), edges_new
        # block.py:375 — every rank needs all source rows: all-gather of the node shards (no-op on one GPU)
        x_full = gather_rows(x, shard_info.nodes if shard_info is not None else None, model_comm_group)
        n_local = x.shape[0]
        csr = Fn.csr_for(edge_index, x_full.shape[0], x_full.shape[0])
        C = self.in_channels
        if group_size(model_comm_group) > 1:
            # local edges only reach local dst rows; aggregate over the full index range and keep our rows (block.py:391)
            agg_full = torch.empty((x_full.shape[0], C), dtype=dt, device=x.device)
            _, edges_new = self.conv.run(x_full, x_full, edge_attr, csr, dt, out=agg_full)
            rank = torch.distributed.get_rank(model_comm_group)
            start = sum(shard_info.nodes[:rank])
            agg_buf = torch.empty((n_local, 2 * C), dtype=dt, device=x.device)
            ops.cast_pad(agg_full[start : start + n_local], dt, out=agg_buf[:, C:])
        else:
            agg_buf = torch.empty((n_local, 2 * C), dtype=dt, device=x.device)
            _, edges_new = self.conv.run(x_full, x_full, edge_attr, csr, dt, out=agg_buf[:, C:])
        return self._node_update(x, agg_buf, dt), edges_new


class GraphConvMapperBlock(GraphConvBaseBlock):
    def forward(
        self,
        x: PairTensor,
        edge_attr: Tensor,
        edge_index: Tensor,
        shard_info: Optional[BipartiteGraphShardInfo] = None,
        model_comm_group=None,
        size=None,
        **layer_kwargs,
    ) -> tuple[PairTensor, Tensor]:
        x_src, x_dst = x
        dt = Fn.compute_dtype(x_src, x_dst, edge_attr)
        if T.wants_grad(self, x_src, x_dst, edge_attr):  # differentiable path (layers/_train.py)
            src_all = x_src
            if group_size(model_comm_group) > 1 and shard_info is not None and shard_info.src_is_sharded():
                src_all = gather_rows_grad(x_src, shard_info.src_nodes, model_comm_group)  # block.py:451
            return T.gnn_block(self, src_all, x_dst, edge_attr, edge_index, dt, bipartite=True, x_src_local=x_src)
        C = self.in_channels
        # sharded (block.py:451-470): x_dst and the edges are this rank's (local dst ids, global src ids); every source row is needed
        src_all = x_src
        if group_size(model_comm_group) > 1 and shard_info is not None and shard_info.src_is_sharded():
            src_all = gather_rows(x_src, shard_info.src_nodes, model_comm_group)
        csr = Fn.csr_for(edge_index, src_all.shape[0], x_dst.shape[0])
        agg_buf = torch.empty((x_dst.shape[0], 2 * C), dtype=dt, device=x_dst.device)
        _, edges_new = self.conv.run(src_all, x_dst, edge_attr, csr, dt, out=agg_buf[:, C:])
        dst_new = self._node_update(x_dst, agg_buf, dt)
        src_new = x_src
        if self.update_src_nodes:  # block.py:475 — the same node_mlp on cat[x_src, x_src]
            src_buf = torch.empty((x_src.shape[0], 2 * C), dtype=dt, device=x_src.device)
            ops.cast_pad(x_src, dt, out=src_buf[:, :C])
            ops.cast_pad(x_src, dt, out=src_buf[:, C:])
            src_new = self.node_mlp.run(src_buf, dt, residual=src_buf[:, :C])
        return (src_new, dst_new), edges_new


# ------------------------------------------------------------------------------------------------------------
# GraphTransformer blocks
# ------------------------------------------------------------------------------------------------------------
class GraphTransformerBaseBlock(Fn.PackOwner):
    """Edge-softmax attention block (block.py:482-687).  Per forward:
    LN -> one GEMM for q|k|v|self -> fused attention(+lin_edge, +self) -> projection GEMM(+skip) -> LN -> MLP GEMMs(+residual).
    """

    def __init__(
        self,
        *,
        in_channels: int,
        hidden_dim: int,
        out_channels: int,
        num_heads: int,
        edge_dim: int,
        bias: bool = True,
        qk_norm: bool = False,
        mlp_implementation: str = "mlp",
        update_src_nodes: bool = False,
        layer_kernels=None,
        attn_channels: Optional[int] = None,
        graph_attention_backend: str = "triton",
        edge_pre_mlp: bool = False,
        **kwargs,
    ) -> None:
        super().__init__()
        k = load_layer_kernels(layer_kernels)
        self.update_src_nodes = update_src_nodes
        self.attn_channels = out_channels if attn_channels is None else attn_channels
        if self.attn_channels <= 0:
            raise ValueError(f"attn_channels must be > 0, got {self.attn_channels}")
        if self.attn_channels % num_heads != 0:
            raise ValueError(f"attn_channels ({self.attn_channels}) must be divisible by num_heads ({num_heads}) in {self.__class__.__name__}.")
        self.out_channels_conv = self.attn_channels // num_heads
        self.num_heads = num_heads
        self.qk_norm = qk_norm
        A = num_heads * self.out_channels_conv
        self.lin_key = k.Linear(in_channels, A)
        self.lin_query = k.Linear(in_channels, A)
        self.lin_value = k.Linear(in_channels, A)
        self.lin_self = k.Linear(in_channels, A, bias=bias)
        self.lin_edge = k.Linear(edge_dim, A)
        self.projection = k.Linear(self.attn_channels, out_channels)
        if self.qk_norm:
            self.q_norm = k.QueryNorm(self.out_channels_conv)
            self.k_norm = k.KeyNorm(self.out_channels_conv)
        self.layer_norm_attention = k.LayerNorm(normalized_shape=in_channels)
        self.layer_norm_mlp_dst = k.LayerNorm(normalized_shape=out_channels)
        self.node_dst_mlp = MLP(out_channels, hidden_dim, out_channels, layer_kernels=k, n_extra_layers=0, layer_norm=False,
                                mlp_implementation=mlp_implementation)  # fmt: skip
        self.edge_pre_mlp = nn.Sequential(k.Linear(edge_dim, edge_dim), k.Activation()) if edge_pre_mlp else nn.Identity()
        if edge_pre_mlp:
            from .mlp import _is_gelu

            if not _is_gelu(self.edge_pre_mlp[1]):  # prepare_edges fuses the activation into the GEMM epilogue as exact-erf GELU
                raise NotImplementedError(f"edge_pre_mlp Activation {type(self.edge_pre_mlp[1]).__name__}: only exact (erf) torch.nn.GELU is fused")
        if graph_attention_backend not in ("triton", "pyg", "b200"):
            raise ValueError(f"Backend '{graph_attention_backend}' not supported for {self.__class__.__name__}")
        # accepted for config compatibility; there is exactly one implementation here (the sm_100a kernel)
        self.graph_attention_backend = graph_attention_backend
        self._pack = Fn.WeightPack()

    # -- lin_edge folding -----------------------------------------------------------------------------------------
    # eproj = W_e a + b_e enters attention linearly, so (include/anemoi_b200.h, form 3)
    #   q.(k + eproj)          = q.k + (W_e,h^T q_h).a + const      -> qw = x_n (W_e,h^T W_q,h)^T : extra columns of the q GEMM
    #   sum alpha (v + eproj)  = sum alpha v + W_e,h abar_h + b_e   -> projection(att + W_e abar) = [att | abar] [W_p | W_p W_e]^T
    # The folded weights are built once per parameter version in fp32.
    def _fold_dims(self) -> tuple[int, int, int]:
        d = self.lin_edge.weight.shape[1]
        dp = (d + 3) // 4 * 4
        return d, dp, Fn.round8(self.num_heads * dp)

    def _use_fold(self, dt: torch.dtype) -> bool:
        return ops.attention_fold_supported(self.attn_channels, self.num_heads, dt, self.lin_edge.weight.shape[1])

    def _w_edge_heads(self) -> Tensor:
        H, Ch = self.num_heads, self.out_channels_conv
        return self.lin_edge.weight.detach().float().view(H, Ch, -1)  # [H, Ch, D]

    def _dst_weight32(self, layers, fold_q: bool) -> tuple[Tensor, Optional[Tensor]]:
        """fp32 (weight, bias) of the dst-side GEMM: rows of ``layers`` concatenated; with ``fold_q`` the rows
        (h, a) = W_e,h[:, a]^T W_q,h (and their bias W_e,h[:, a]^T b_q,h) are appended."""
        w, b = Fn.cat_linear32(layers)
        if not fold_q:
            return w, b
        H, Ch = self.num_heads, self.out_channels_conv
        d, dp, hdp = self._fold_dims()
        we = self._w_edge_heads()
        wq = self.lin_query.weight.detach().float().view(H, Ch, -1)
        wf = torch.zeros(hdp, wq.shape[-1], device=wq.device)
        wf[: H * dp].view(H, dp, -1)[:, :d] = torch.einsum("hca,hci->hai", we, wq)
        bf = torch.zeros(hdp, device=we.device)
        if self.lin_query.bias is not None:
            bf[: H * dp].view(H, dp)[:, :d] = torch.einsum("hca,hc->ha", we, self.lin_query.bias.detach().float().view(H, Ch))
        if b is None:
            b = torch.zeros(w.shape[0], device=w.device)
        return torch.cat([w, wf], 0), torch.cat([b, bf])

    def _dst_gemm(self, x_raw: Tensor, ln: nn.Module, layers, dt: torch.dtype, cond: Optional[Tensor] = None) -> Tensor:
        """LayerNorm + dst-side GEMM (q | ... | self | qw); the LayerNorm is folded into the GEMM on the bf16 path."""
        fold_q = self._use_fold(dt) and not self.qk_norm
        key = ("dst", tuple(id(l) for l in layers), fold_q)
        srcs = Fn.linear_sources(layers) + [self.lin_edge.weight]
        return Fn.ln_linear(self._pack, x_raw, ln, key, srcs, lambda: self._dst_weight32(layers, fold_q), dt, cond=cond)

    def _qw_blockdiag(self, dt: torch.dtype) -> Tensor:
        """[hdp, A] block-diagonal W_e^T for the qk_norm case (qw must be taken from the normalised query)."""
        H, Ch = self.num_heads, self.out_channels_conv
        d, dp, hdp = self._fold_dims()

        def build() -> Tensor:
            we = self._w_edge_heads()
            w = torch.zeros(hdp, H * Ch, device=we.device)
            for h in range(H):
                w[h * dp : h * dp + d, h * Ch : (h + 1) * Ch] = we[h].t()
            return w.to(dt).contiguous()

        return self._pack.get(("w_qw_bd", dt), [self.lin_edge.weight], build)

    def _proj_weight(self, dt: torch.dtype, fold: bool) -> Tensor:
        if not fold:
            return self._pack.weight([self.projection], dt)
        H, Ch = self.num_heads, self.out_channels_conv
        d, dp, hdp = self._fold_dims()

        def build() -> Tensor:
            we = self._w_edge_heads()
            wp = self.projection.weight.detach().float()  # [C_out, A]
            wf = torch.zeros(wp.shape[0], hdp, device=wp.device)
            wf[:, : H * dp].view(-1, H, dp)[:, :, :d] = torch.einsum("ohc,hca->oha", wp.view(-1, H, Ch), we)
            return torch.cat([wp, wf], 1).to(dt).contiguous()

        return self._pack.get(("w_proj_fold", dt), [self.projection.weight, self.lin_edge.weight], build)

    # -- pieces -------------------------------------------------------------------------------------------------
    def prepare_edges(self, edge_attr: Tensor, dt: torch.dtype = torch.bfloat16) -> Tensor:
        """Raw edge attributes -> fp32 zero-padded operand of the fused lin_edge (after edge_pre_mlp if present):
        16 floats per edge for the folded slab kernel, ceil4(d_e) for the in-kernel projection path."""
        if not isinstance(self.edge_pre_mlp, nn.Identity):
            lin = self.edge_pre_mlp[0]
            edge_attr = Fn.fused_linear(self._pack, edge_attr.float() if edge_attr.dtype != torch.float32 else edge_attr, [lin], torch.float32,
                                        gelu=True)  # fmt: skip
        return Fn.pad_edge_attr(edge_attr, ops.ATTN_MAX_EDGE_DIM if self._use_fold(dt) else 0)

    def _attend_project(self, x_dst: Tensor, ln_dst: nn.Module, k: Tensor, v: Tensor, dst_layers, edge_attr_p: Tensor, csr: ops.GraphCSR,
                        x_skip: Tensor, dt: torch.dtype, dst_buf: Optional[Tensor] = None, want_stats: bool = False,
                        k_prenormed: bool = False, cond: Optional[Tensor] = None) -> Tensor:  # fmt: skip
        """dst-side GEMM (q | [k | v |] self | qw) -> attention (+ self) -> projection (+ skip) -> LN -> MLP (+ residual).

        ``dst_layers`` = the Linear containers of the dst-side GEMM *after* lin_query (processor: key, value, self — k and v then
        come out of the same GEMM and ``k``/``v`` are None; mapper: self only).  ``dst_buf`` lets the processor pass a precomputed GEMM.
        """
        A, H = self.attn_channels, self.num_heads
        fold = self._use_fold(dt)
        d, dp, hdp = self._fold_dims()
        layers = [self.lin_query] + list(dst_layers)
        if dst_buf is None:
            dst_buf = self._dst_gemm(x_dst, ln_dst, layers, dt, cond=cond)
        n_lin = len(layers)
        q = dst_buf[:, :A]
        x_r = dst_buf[:, (n_lin - 1) * A : n_lin * A]
        if k is None:
            k, v = dst_buf[:, A : 2 * A], dst_buf[:, 2 * A : 3 * A]
        if self.qk_norm:
            for t, norm in ((q, self.q_norm),) + (() if k_prenormed else ((k, self.k_norm),)):
                ops.layer_norm(t, self._pack.f32(norm.weight), self._pack.f32(getattr(norm, "bias", None)), norm.eps, out=t, groups=H)
        b_e = self._pack.f32(self.lin_edge.bias)
        # Serpentine schedule (ops.set_traversal): q | k | v were just written top-down by the GEMM, so the attention walks the rows bottom-up
        # and finds the last-written half of the 183 MB buffer in L2; the projection then runs top-down again over what attention wrote last.
        ops.set_traversal(gemm=SERPENTINE)
        if fold:
            qw = ops.linear(q, self._qw_blockdiag(dt)) if self.qk_norm else dst_buf[:, n_lin * A :]
            att = torch.empty((q.shape[0], A + hdp), dtype=dt, device=q.device)
            if hdp != H * dp:
                att[:, A + H * dp :].zero_()
            ops.gt_attention(q, k, v, csr, H, edge_attr=edge_attr_p, b_edge=b_e, qw=qw, abar=att[:, A:], dp=dp, add=x_r, out=att[:, :A],
                             tiles=attention_tile_plan(csr, A, H, dt, dp))  # fmt: skip
        else:
            w_e = self._pack.get(("w_edge",), [self.lin_edge.weight], lambda: self.lin_edge.weight.detach().float().contiguous())
            att = ops.gt_attention(q, k, v, csr, H, edge_attr=edge_attr_p, w_edge=w_e, b_edge=b_e, add=x_r)
        ops.set_traversal()
        return self._project_mlp(att, x_skip, dt, fold, want_stats, cond)

    def _heads_attention(self, q: Tensor, qw: Tensor, k: Tensor, v: Tensor, x_r: Tensor, dst_sizes: list, src_sizes: Optional[list], ea: Tensor,
                         edge_index: Tensor, group, dt: torch.dtype) -> Tensor:  # fmt: skip
        """Attention of the "heads" (Ulysses) strategy.  ``q | qw | x_r`` are this rank's dst rows, ``k | v`` its src rows (``src_sizes``) or
        ALL src rows (``src_sizes is None``).  The rows are exchanged so that each rank holds all nodes for its H / P heads (one all-to-all),
        attention runs over the FULL edge list for those heads, a second all-to-all brings every rank its own rows for all heads; returns
        ``[att + self | abar]`` for the projection GEMM.  qk_norm is applied by the row owners before the exchange."""
        from ..distributed.graph import SegmentedCapture
        from ..distributed.graph import _exchange
        from ..distributed.graph import group_rank

        def xchg(send: Tensor, in_splits: list, out_splits: list) -> Tensor:
            if SegmentedCapture.active is None:
                return _exchange(send, in_splits, out_splits, group)
            out = send.new_empty((sum(out_splits), send.shape[1]))  # static buffer of the capture pool; the collective runs between two segments
            return SegmentedCapture.active.exchange(send, in_splits, out_splits, group, out)

        A, H, Ch = self.attn_channels, self.num_heads, self.out_channels_conv
        P, me = group_size(group), group_rank(group)
        n_l, N_dst = q.shape[0], sum(dst_sizes)
        if H % P:
            raise ValueError(f"heads strategy: num_heads ({H}) must be divisible by the model group size ({P})")
        if not self._use_fold(dt) or not ops.attention_fold_supported(A // P, H // P, dt, self.lin_edge.weight.shape[1]):
            raise NotImplementedError("heads strategy: implemented for shapes the folded lin_edge attention kernel handles")
        Hl = H // P
        d, dp, hdp = self._fold_dims()
        if (Hl * (Ch + dp) * torch.empty(0, dtype=dt).element_size()) % 16:
            raise NotImplementedError(f"heads strategy: the exchanged rows [q | qw] of {Hl} heads x ({Ch} + {dp}) {dt} are not multiples of 16 bytes; "
                                      "choose num_heads / model-parallel size so that (H / P) * (Ch + dp) * elemsize % 16 == 0")  # fmt: skip
        if self.qk_norm:
            for t, norm in ((q, self.q_norm), (k, self.k_norm)):
                ops.layer_norm(t, self._pack.f32(norm.weight), self._pack.f32(getattr(norm, "bias", None)), norm.eps, out=t, groups=H)
            qw = ops.linear(q, self._qw_blockdiag(dt))[:, : H * dp]

        def to_heads(parts: list, n_rows: int, sizes: list) -> Tensor:
            """[n_rows, H * w_i] column blocks -> every rank receives ALL rows (global order) of its head group: [sum(sizes), Hl * sum(w_i)]"""
            send = torch.cat([t.reshape(n_rows, P, -1) for t in parts], dim=2).permute(1, 0, 2)
            return xchg(send.reshape(P * n_rows, -1).contiguous(), [n_rows] * P, sizes)

        qa = to_heads([q, qw], n_l, dst_sizes)  # [N_dst, Hl * (Ch + dp)] = q | qw of my heads
        if src_sizes is None:  # sources replicated: my heads are column slices, nothing travels
            ka, va = k[:, me * Hl * Ch : (me + 1) * Hl * Ch], v[:, me * Hl * Ch : (me + 1) * Hl * Ch]
        else:
            kva = to_heads([k, v], k.shape[0], src_sizes)
            ka, va = kva[:, : Hl * Ch], kva[:, Hl * Ch :]
        csr = Fn.csr_for(edge_index, ka.shape[0], N_dst)
        out_h = torch.empty((N_dst, Hl * (Ch + dp)), dtype=dt, device=q.device)
        b_e = self._pack.f32(self.lin_edge.bias)
        ops.gt_attention(qa[:, : Hl * Ch], ka, va, csr, Hl, edge_attr=ea, b_edge=None if b_e is None else b_e[me * Hl * Ch : (me + 1) * Hl * Ch],
                         qw=qa[:, Hl * Ch :], abar=out_h[:, Hl * Ch :], dp=dp, out=out_h[:, : Hl * Ch])  # fmt: skip
        back = xchg(out_h, dst_sizes, [n_l] * P).reshape(P, n_l, Hl * (Ch + dp))  # [source rank = head group, local row, att | abar]
        att = torch.empty((n_l, A + hdp), dtype=dt, device=q.device)
        if hdp != H * dp:
            att[:, A + H * dp :].zero_()
        att_h = back[:, :, : Hl * Ch].permute(1, 0, 2).reshape(n_l, A)
        ops.cast_pad(ops.add(att_h.contiguous(), x_r), dt, out=att[:, :A])  # + self term (fused into the kernel in the other strategies)
        ops.cast_pad(back[:, :, Hl * Ch :].permute(1, 0, 2).reshape(n_l, H * dp).contiguous(), dt, out=att[:, A : A + H * dp])
        return att

    def _project_mlp(self, att: Tensor, x_skip: Tensor, dt: torch.dtype, fold: bool, want_stats: bool, cond: Optional[Tensor]) -> Tensor:
        """projection (+ skip) -> LayerNorm -> MLP (+ residual) on the attention output ``att`` (``[att + self | abar]`` in the folded form)."""
        skip = x_skip if x_skip.dtype in Fn.SUPPORTED else x_skip.float()
        wp = self._proj_weight(dt, fold)
        # the projection epilogue also produces the row statistics of its output for layer_norm_mlp_dst (folded into the MLP's first GEMM),
        # and the MLP's last GEMM those of the block output for the next block's layer_norm_attention
        out = Fn.linear_with_stats(Fn.as_operand(att, dt, wp.shape[1]), wp, self._pack.bias([self.projection]), residual=skip, want_stats=True)
        return self.node_dst_mlp.run(out, dt, residual=out, pre_ln=self.layer_norm_mlp_dst, want_stats=want_stats, cond=cond, serpentine=SERPENTINE)


class GraphTransformerProcessorBlock(GraphTransformerBaseBlock):
    def __init__(self, *, shard_strategy: str = "edges", **kwargs) -> None:
        super().__init__(**kwargs)
        if shard_strategy not in ("edges", "heads"):
            raise ValueError(f"Invalid shard strategy '{shard_strategy}'")
        self.shard_strategy = shard_strategy

    def forward(
        self,
        x: Tensor,
        edge_attr: Tensor,
        edge_index: Tensor,
        shard_info: Optional[GraphShardInfo] = None,
        batch_size: int = 1,
        size=None,
        model_comm_group=None,
        cond: Optional[Tensor] = None,
        edges_are_dst_sorted: bool = True,
        edge_attr_prepared: Optional[Tensor] = None,
        **kwargs,
    ) -> tuple[Tensor, Tensor]:
        dt = Fn.compute_dtype(x)
        edge_attr_in = edge_attr
        if not edges_are_dst_sorted:  # the reference sorts here (edge_index_to_csc(..., edges_are_dst_sorted=False), block.py:779-782)
            from ..distributed.khop_edges import ensure_edges_are_dst_sorted

            edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, False)
            edge_attr_prepared = None
        if T.wants_grad(self, x, edge_attr):  # differentiable path (layers/_train.py)
            plan = heads = None
            if group_size(model_comm_group) > 1:
                if self.shard_strategy == "heads":
                    heads = (model_comm_group, list(shard_info.nodes), list(shard_info.nodes))
                else:
                    plan = halo_plan_for(edge_index, shard_info.nodes, model_comm_group)  # same plan as the inference path (block.py:1120-1183)
            return T.gt_block(self, None, x, edge_attr, edge_index, dt, None, cond, plan=plan, heads=heads), edge_attr_in
        A = self.attn_channels
        ln = self.layer_norm_attention  # with a ConditionalLayerNorm kernel both LayerNorms of the block take ``cond`` (block.py:1233-1271)
        dst_layers = [self.lin_key, self.lin_value, self.lin_self]
        ea = edge_attr_prepared if edge_attr_prepared is not None else self.prepare_edges(edge_attr, dt)
        world = group_size(model_comm_group)
        if world == 1:
            csr = Fn.csr_for(edge_index, x.shape[0], x.shape[0])
            return self._attend_project(x, ln, None, None, dst_layers, ea, csr, x, dt, want_stats=True, cond=cond), edge_attr
        if self.shard_strategy == "heads":
            return self._forward_heads(x, ln, ea, edge_index, shard_info, model_comm_group, dt, cond), edge_attr
        # edges strategy (block.py:1120-1183): each rank owns a dst range and needs the k | v rows of the source nodes its edges name
        if self.qk_norm and not HALO_EXCHANGE:
            raise NotImplementedError("qk_norm with the all-gather form of the sharded processor (use the halo exchange)")
        if not HALO_EXCHANGE:  # A/B switch: all-gather every k | v row (round-1 first version)
            buf = self._dst_gemm(x, ln, [self.lin_query] + dst_layers, dt, cond=cond)
            kv_full = gather_rows(buf[:, A : 3 * A], shard_info.nodes, model_comm_group)
            csr = Fn.csr_for(edge_index, kv_full.shape[0], x.shape[0])
            out = self._attend_project(x, ln, kv_full[:, :A], kv_full[:, A:], dst_layers, ea, csr, x, dt, dst_buf=buf, want_stats=True, cond=cond)
            return out, edge_attr
        # halo exchange (distributed/halo.py): the k | v GEMM writes this rank's rows into the head of a compact table, one gather kernel
        # packs the rows the other ranks asked for, one all-to-all drops the rows we need into the tail; the edge list was relabelled
        # onto the table once.  q | self | qw come from a second GEMM on the same (tagged) LayerNorm statistics.
        plan = halo_plan_for(edge_index, shard_info.nodes, model_comm_group)
        table = plan.table(2 * A, dt, x.device)  # symmetric peer buffer (one per plan, re-used by every layer) or a fresh tensor
        kv_layers = [self.lin_key, self.lin_value]
        Fn.ln_linear(self._pack, x, ln, ("kv", id(self.lin_key), id(self.lin_value)), Fn.linear_sources(kv_layers),
                     lambda: Fn.cat_linear32(kv_layers), dt, cond=cond, out=table[: plan.n_local])  # fmt: skip
        if self.qk_norm:  # per-row, per-head: commutes with the exchange, so every rank normalises the keys it owns once
            ops.layer_norm(table[: plan.n_local, :A], self._pack.f32(self.k_norm.weight), self._pack.f32(getattr(self.k_norm, "bias", None)),
                           self.k_norm.eps, out=table[: plan.n_local, :A], groups=self.num_heads)  # fmt: skip
        plan.exchange_start(table)
        buf = self._dst_gemm(x, ln, [self.lin_query, self.lin_self], dt, cond=cond)  # runs while the halo rows travel
        plan.exchange_finish()
        csr = Fn.csr_for(plan.edge_index, plan.n_table, x.shape[0])
        out = self._attend_project(x, ln, table[:, :A], table[:, A:], [self.lin_self], ea, csr, x, dt, dst_buf=buf, want_stats=True,
                                   k_prenormed=True, cond=cond)  # fmt: skip
        return out, edge_attr


    def _forward_heads(self, x: Tensor, ln: nn.Module, ea: Tensor, edge_index: Tensor, shard_info, group, dt: torch.dtype, cond) -> Tensor:
        """"heads" (Ulysses) strategy (block.py:689-759, 1185-1217): nodes are sharded outside the attention, heads inside it
        (``_heads_attention``).  ``edge_index`` / ``ea`` are the FULL dst-sorted edge list and prepared attributes (the processor does not
        shard edges in this mode)."""
        A, H = self.attn_channels, self.num_heads
        d, dp, hdp = self._fold_dims()
        buf = self._dst_gemm(x, ln, [self.lin_query, self.lin_key, self.lin_value, self.lin_self], dt, cond=cond)  # q | k | v | self | qw
        q, k, v, x_r = buf[:, :A], buf[:, A : 2 * A], buf[:, 2 * A : 3 * A], buf[:, 3 * A : 4 * A]
        sizes = list(shard_info.nodes)
        att = self._heads_attention(q, buf[:, 4 * A : 4 * A + H * dp], k, v, x_r, sizes, sizes, ea, edge_index, group, dt)
        return self._project_mlp(att, x, dt, True, True, cond)


class GraphTransformerMapperBlock(GraphTransformerBaseBlock):
    def __init__(self, *, shard_strategy: str = "edges", **kwargs) -> None:
        super().__init__(**kwargs)
        k = load_layer_kernels(kwargs.get("layer_kernels"))
        in_channels, out_channels = kwargs["in_channels"], kwargs["out_channels"]
        self.layer_norm_attention_src = k.LayerNorm(normalized_shape=in_channels)
        self.layer_norm_attention_dest = self.layer_norm_attention  # alias, as in the reference (block.py:941)
        if self.update_src_nodes:
            self.layer_norm_mlp_src = k.LayerNorm(normalized_shape=out_channels)
            self.node_src_mlp = MLP(out_channels, kwargs["hidden_dim"], out_channels, layer_kernels=k, n_extra_layers=0, layer_norm=False,
                                    mlp_implementation=kwargs.get("mlp_implementation", "mlp"))  # fmt: skip
        else:
            self.layer_norm_mlp_src = nn.Identity()
            self.node_src_mlp = nn.Identity()
        self.shard_strategy = shard_strategy

    def forward(
        self,
        x: PairTensor,
        edge_attr: Tensor,
        edge_index: Tensor,
        shard_info: Optional[BipartiteGraphShardInfo] = None,
        batch_size: int = 1,
        size=None,
        model_comm_group=None,
        cond=None,
        edges_are_dst_sorted: bool = True,
        **layer_kwargs,
    ) -> tuple[PairTensor, Tensor]:
        x_src, x_dst = x
        cond_src, cond_dst = cond if cond is not None else (None, None)  # (block.py:978-980)
        dt = Fn.compute_dtype(x_src, x_dst)
        if T.wants_grad(self, x_src, x_dst, edge_attr):  # differentiable path (layers/_train.py)
            plan = heads = None
            if group_size(model_comm_group) > 1:
                if self.shard_strategy == "heads":
                    heads = (model_comm_group, list(shard_info.dst_nodes), list(shard_info.src_nodes) if shard_info.src_is_sharded() else None)
                elif shard_info is not None and shard_info.src_is_sharded():
                    plan = halo_plan_for(edge_index, shard_info.src_nodes, model_comm_group)
            dst_new = T.gt_block(self, x_src, x_dst, edge_attr, edge_index, dt, self.layer_norm_attention_src, cond, plan=plan, heads=heads)
            src_new = x_src
            if self.update_src_nodes:
                src_new = T.mlp(self.node_src_mlp, x_src.to(dt), dt, residual=x_src, pre_ln=self.layer_norm_mlp_src, cond=cond_src)
            return (src_new, dst_new), edge_attr
        A = self.attn_channels
        if not edges_are_dst_sorted:  # the reference sorts here (block.py:779-782)
            from ..distributed.khop_edges import ensure_edges_are_dst_sorted

            edge_attr, edge_index = ensure_edges_are_dst_sorted(edge_attr, edge_index, False)
        kv_layers = [self.lin_key, self.lin_value]
        world = group_size(model_comm_group)
        halo = world > 1 and self.shard_strategy != "heads" and shard_info is not None and shard_info.src_is_sharded()
        plan = table = None
        if halo:  # the k | v GEMM writes this rank's rows straight into the head of the halo table (see below)
            plan = halo_plan_for(edge_index, shard_info.src_nodes, model_comm_group)
            table = plan.table(2 * A, dt, x_src.device)
        kv = Fn.ln_linear(self._pack, x_src, self.layer_norm_attention_src, ("kv",), Fn.linear_sources(kv_layers), lambda: Fn.cat_linear32(kv_layers), dt,
                          cond=cond_src, **({"out": table[: plan.n_local]} if halo else {}))  # fmt: skip
        if group_size(model_comm_group) > 1 and self.shard_strategy == "heads":
            # heads strategy (mapper.py:388-478): full edge list, heads sharded inside the attention
            d_, dp_, _ = self._fold_dims()
            buf = self._dst_gemm(x_dst, self.layer_norm_attention_dest, [self.lin_query, self.lin_self], dt, cond=cond_dst)  # q | self | qw
            att = self._heads_attention(buf[:, :A], buf[:, 2 * A : 2 * A + self.num_heads * dp_], kv[:, :A], kv[:, A:], buf[:, A : 2 * A],
                                        list(shard_info.dst_nodes), list(shard_info.src_nodes) if shard_info.src_is_sharded() else None,
                                        self.prepare_edges(edge_attr, dt), edge_index, model_comm_group, dt)  # fmt: skip
            dst_new = self._project_mlp(att, x_dst, dt, True, False, cond_dst)
            src_new = x_src
            if self.update_src_nodes:
                src_new = self.node_src_mlp.run(x_src, dt, residual=x_src if x_src.dtype in Fn.SUPPORTED else x_src.float(),
                                                pre_ln=self.layer_norm_mlp_src, cond=cond_src)  # fmt: skip
            return (src_new, dst_new), edge_attr
        if halo:
            # edges strategy with sharded sources (reference mapper.py:248-297 / khop_edges.py:317-409 all-gathers the source rows and drops
            # the unconnected ones): only the k | v rows this rank's edges name travel, through the same halo plan as the processor's, over
            # SOURCE rows partitioned by shard_info.src_nodes; the dst-side GEMM runs while they do.
            if self.qk_norm:  # per row and head: every rank normalises the keys it owns once, before they travel
                ops.layer_norm(table[: plan.n_local, :A], self._pack.f32(self.k_norm.weight), self._pack.f32(getattr(self.k_norm, "bias", None)),
                               self.k_norm.eps, out=table[: plan.n_local, :A], groups=self.num_heads)  # fmt: skip
            plan.exchange_start(table)
            buf = self._dst_gemm(x_dst, self.layer_norm_attention_dest, [self.lin_query, self.lin_self], dt, cond=cond_dst)
            plan.exchange_finish()
            csr = Fn.csr_for(plan.edge_index, plan.n_table, x_dst.shape[0])
            dst_new = self._attend_project(x_dst, self.layer_norm_attention_dest, table[:, :A], table[:, A:], [self.lin_self],
                                           self.prepare_edges(edge_attr, dt), csr, x_dst, dt, dst_buf=buf, k_prenormed=True, cond=cond_dst)  # fmt: skip
        else:
            csr = Fn.csr_for(edge_index, kv.shape[0], x_dst.shape[0])
            dst_new = self._attend_project(x_dst, self.layer_norm_attention_dest, kv[:, :A], kv[:, A:], [self.lin_self], self.prepare_edges(edge_attr, dt),
                                           csr, x_dst, dt, cond=cond_dst)  # fmt: skip
        src_new = x_src
        if self.update_src_nodes:
            src_new = self.node_src_mlp.run(x_src, dt, residual=x_src if x_src.dtype in Fn.SUPPORTED else x_src.float(),
                                            pre_ln=self.layer_norm_mlp_src, cond=cond_src)
        return (src_new, dst_new), edge_attr
