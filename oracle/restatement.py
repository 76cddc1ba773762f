"""oracle/restatement.py — CPU restatement of the anemoi-models graph message-passing forward.

*** TEST INFRASTRUCTURE.  Not shipped, not imported by the product package. ***
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module, and only as the checker / reported baseline.

It is a plain-PyTorch (fp32, CPU) functional restatement of the reference algorithm, driven by a
reference ``state_dict`` (same parameter names as the reference modules), with no dependency on
``torch_geometric`` / ``hydra`` / ``anemoi.utils``.  Every function cites the reference file:line it
follows (paths relative to ``/root/reference/models/src/anemoi/models``).

Pinning (see DESIGN.md "Oracle"): the reference ships NO literal golden vectors for this path
(SURVEY.md §8c).  The restatement is therefore pinned against outputs of the *unmodified reference
modules* run in the build container through ``oracle/standins`` (``oracle/gen_golden.py`` →
``tests/golden/*.pt``); ``tests/test_oracle_golden.py`` checks restatement == golden on every run.
Third-party arithmetic (PyG ``scatter``/``softmax``/``index2ptr``; pinned only as
``torch-geometric>=2.3`` in ``models/pyproject.toml:44``) is restated from its published semantics
and additionally cross-checked against the reference's independent statement of the same maths,
the Triton kernel ``triton/gt.py:81-179`` (online softmax), in ``gt_attention_online``.
"""

from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

StateDict = dict


# ----------------------------------------------------------------------------------------------
# integer path
# ----------------------------------------------------------------------------------------------
def get_balanced_partition_sizes(total_size: int, n_partitions: int) -> list[int]:
    """distributed/balanced_partition.py:16-41 — first ``rem`` parts get one extra element."""
    base, rem = divmod(total_size, n_partitions)
    return [base + 1] * rem + [base] * (n_partitions - rem)


def sort_edge_index_by_dst(edge_index: Tensor) -> tuple[Tensor, Tensor]:
    """distributed/khop_edges.py:37-40 — stable sort of edge_index by destination (row 1)."""
    perm = torch.sort(edge_index[1], stable=True)[1]
    return edge_index[:, perm], perm


def is_edge_index_dst_sorted(edge_index: Tensor) -> bool:
    """distributed/khop_edges.py:43-48."""
    dst = edge_index[1]
    return True if dst.numel() <= 1 else bool(torch.all(dst[1:] >= dst[:-1]).item())


def index2ptr(index: Tensor, size: int) -> Tensor:
    """torch_geometric.utils.sparse.index2ptr: ptr[i] = #{index < i}; int64 for int64 input."""
    counts = torch.bincount(index, minlength=size)
    ptr = torch.zeros(size + 1, dtype=index.dtype)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr


def edge_index_to_csc(edge_index: Tensor, num_nodes: tuple[int, int], edges_are_dst_sorted: bool = False):
    """triton/utils.py:25-70 — (row, colptr), perm, (rowptr, edge_id_per_src, edge_dst)."""
    perm = None
    if not edges_are_dst_sorted:
        edge_index, perm = sort_edge_index_by_dst(edge_index)
    row, col = edge_index[0], edge_index[1]
    colptr = index2ptr(col, num_nodes[1])
    row_sorted = torch.sort(row)[0]
    rowptr = index2ptr(row_sorted, num_nodes[0])
    edge_id_per_src = torch.argsort(row, stable=True)
    return (row, colptr), perm, (rowptr, edge_id_per_src, col)


def build_graph_partition(edge_index: Tensor, num_parts: int, num_nodes: tuple[int, int]):
    """distributed/khop_edges.py:154-189 — balanced dst splits and per-part edge counts."""
    n_dst = num_nodes[1]
    dst_splits = get_balanced_partition_sizes(n_dst, num_parts)
    deg = torch.bincount(edge_index[1], minlength=n_dst)
    edge_splits = [int(c.sum()) for c in torch.split(deg, dst_splits)]
    return dst_splits, edge_splits


def drop_unconnected_src_nodes(n_src: int, edge_index: Tensor) -> tuple[Tensor, Tensor]:
    """distributed/khop_edges.py:474-500 — connected src ids (sorted unique) and relabelled edge_index."""
    edge_index = edge_index.clone()
    connected = torch.unique(edge_index[0])
    relabel = torch.empty(n_src, dtype=torch.long)
    relabel[connected] = torch.arange(connected.numel())
    edge_index[0] = relabel[edge_index[0]]
    return connected, edge_index


def materialise_chunk(dst_splits, edge_splits, part: int, n_src: int, edge_index: Tensor):
    """distributed/khop_edges.py:78-132 — (dst_range, edge_range, connected_src, relabelled edge_index)."""
    e0 = sum(edge_splits[:part])
    e1 = e0 + edge_splits[part]
    d0 = sum(dst_splits[:part])
    d1 = d0 + dst_splits[part]
    ei = edge_index[:, e0:e1].clone()
    ei[1] -= d0
    connected, ei = drop_unconnected_src_nodes(n_src, ei)
    return (d0, d1), (e0, e1), connected, ei


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def _linear(sd: StateDict, name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _layer_norm(sd: StateDict, name: str, x: Tensor, autocast_ln: bool = False, cond: Optional[Tensor] = None) -> Tensor:
    """torch.nn.LayerNorm (eps 1e-5); AutocastLayerNorm casts back to x.dtype (normalization.py:19-31); ConditionalLayerNorm
    (normalization.py:34-94): LN without affine, times (1 + scale(cond)), plus bias(cond)."""
    if name + ".scale.weight" in sd:
        y = F.layer_norm(x, (x.shape[-1],), None, None, 1e-5)
        return (y * (_linear(sd, name + ".scale", cond) + 1.0) + _linear(sd, name + ".bias", cond)).type_as(x)
    w = sd[name + ".weight"]
    y = F.layer_norm(x, (w.shape[0],), w, sd.get(name + ".bias"), 1e-5)
    return y.type_as(x) if autocast_ln else y


GATING = ["swiglu"]  # the gating activation has no parameters, so the state_dict cannot name it: set with ``gated_mlp(kind)``


class gated_mlp:
    """``with gated_mlp("geglu"): ...`` — which ``mlp_implementation`` the GatedMLPLayer entries of the state_dict were built with."""

    def __init__(self, kind: str):
        self.kind, self.prev = kind, None

    def __enter__(self):
        self.prev, GATING[0] = GATING[0], self.kind

    def __exit__(self, *exc):
        GATING[0] = self.prev


def mlp(sd: StateDict, prefix: str, x: Tensor, autocast_ln: bool = False) -> Tensor:
    """layers/mlp.py:97-179 — (Linear, GELU(erf))×k, Linear, optional LayerNorm; k from the state_dict."""
    idx = sorted({int(k[len(prefix) + 5 :].split(".")[0]) for k in sd if k.startswith(prefix + ".mlp.")})
    for n, i in enumerate(idx):
        if f"{prefix}.mlp.{i}.gate_proj.weight" in sd:  # GatedMLPLayer (mlp.py:38-53): gating(gate_proj(x)) * value_proj(x)
            gate = {"glu": torch.sigmoid, "swiglu": F.silu, "geglu": F.gelu, "reglu": F.relu}[GATING[0]]
            x = gate(_linear(sd, f"{prefix}.mlp.{i}.gate_proj", x)) * _linear(sd, f"{prefix}.mlp.{i}.value_proj", x)
            continue
        x = _linear(sd, f"{prefix}.mlp.{i}", x)
        if n < len(idx) - 1:
            x = F.gelu(x)
    if prefix + ".layer_norm.weight" in sd:
        x = _layer_norm(sd, prefix + ".layer_norm", x, autocast_ln)
    return x


def scatter_sum(src: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_geometric.utils.scatter(reduce='sum') as used at conv.py:79."""
    out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


# ----------------------------------------------------------------------------------------------
# GraphConv (GNN)
# ----------------------------------------------------------------------------------------------
def graph_conv(sd, prefix, x_src, x_dst, edge_attr, edge_index, autocast_ln=False):
    """layers/conv.py:66-81 — e' = edge_mlp(cat[x_i, x_j, e]) + e ; out = scatter_sum(e', dst)."""
    x_i = x_dst.index_select(0, edge_index[1])
    x_j = x_src.index_select(0, edge_index[0])
    edges_new = mlp(sd, prefix + ".edge_mlp", torch.cat([x_i, x_j, edge_attr], dim=1), autocast_ln) + edge_attr
    out = scatter_sum(edges_new, edge_index[1], x_dst.shape[0])
    return out, edges_new


def gnn_processor_block(sd, prefix, x, edge_attr, edge_index, autocast_ln=False):
    """layers/block.py:362-395 (single device: sync_tensor / shard_tensor are identities)."""
    if any(k.startswith(prefix + ".emb_edges.mlp.0.") for k in sd):  # Linear (.weight) or GatedMLPLayer (.gate_proj.weight)
        edge_attr = mlp(sd, prefix + ".emb_edges", edge_attr, autocast_ln)
    out, edges_new = graph_conv(sd, prefix + ".conv", x, x, edge_attr, edge_index, autocast_ln)
    nodes_new = mlp(sd, prefix + ".node_mlp", torch.cat([x, out], dim=1), autocast_ln) + x
    return nodes_new, edges_new


def gnn_mapper_block(sd, prefix, x_src, x_dst, edge_attr, edge_index, update_src_nodes, autocast_ln=False):
    """layers/block.py:441-479 — bipartite; forward mapper re-uses node_mlp on cat[x_src, x_src] (:475)."""
    out, edges_new = graph_conv(sd, prefix + ".conv", x_src, x_dst, edge_attr, edge_index, autocast_ln)
    dst_new = mlp(sd, prefix + ".node_mlp", torch.cat([x_dst, out], dim=1), autocast_ln) + x_dst
    src_new = x_src
    if update_src_nodes:
        src_new = mlp(sd, prefix + ".node_mlp", torch.cat([x_src, x_src], dim=1), autocast_ln) + x_src
    return (src_new, dst_new), edges_new


def gnn_processor(sd, x, edge_attr, edge_index, num_layers, autocast_ln=False, max_layers=None):
    """layers/processor.py:397-455 — edges_new of layer l is edge_attr of layer l+1."""
    for layer in range(num_layers if max_layers is None else min(num_layers, max_layers)):
        x, edge_attr = gnn_processor_block(sd, f"proc.{layer}", x, edge_attr, edge_index, autocast_ln)
    return x


def gnn_forward_mapper(sd, x_src, x_dst, edge_attr, edge_index, autocast_ln=False):
    """layers/mapper.py:774-834, 863-965 — returns (updated src embedding, dst)."""
    e = mlp(sd, "emb_edges", edge_attr, autocast_ln)
    xs = mlp(sd, "emb_nodes_src", x_src, autocast_ln)
    xd = mlp(sd, "emb_nodes_dst", x_dst, autocast_ln)
    (xs, xd), _ = gnn_mapper_block(sd, "proc", xs, xd, e, edge_index, True, autocast_ln)
    return xs, xd


def gnn_backward_mapper(sd, x_src, x_dst, edge_attr, edge_index, autocast_ln=False):
    """layers/mapper.py:774-834, 968-1087 — pre_process identity, node_data_extractor MLP without LN."""
    e = mlp(sd, "emb_edges", edge_attr, autocast_ln)
    (_, xd), _ = gnn_mapper_block(sd, "proc", x_src, x_dst, e, edge_index, False, autocast_ln)
    return mlp(sd, "node_data_extractor", xd, autocast_ln)


# ----------------------------------------------------------------------------------------------
# GraphTransformer
# ----------------------------------------------------------------------------------------------
def gt_attention(q: Tensor, k: Tensor, v: Tensor, e: Tensor, edge_index: Tensor, n_dst: int) -> Tensor:
    """layers/conv.py:103-147 with PyG softmax: max-subtracted, denominator + 1e-16. q,k,v,e: [*, H, Ch]."""
    src, dst = edge_index[0], edge_index[1]
    ch = q.shape[-1]
    k_j = k.index_select(0, src) + e
    alpha = (q.index_select(0, dst) * k_j).sum(dim=-1) / ch**0.5  # [E, H]
    idx = dst.view(-1, 1).expand_as(alpha)
    amax = alpha.new_zeros((n_dst, alpha.shape[1])).scatter_reduce_(0, idx, alpha, "amax", include_self=False)
    ex = (alpha - amax.index_select(0, dst)).exp()
    den = scatter_sum(ex, dst, n_dst) + 1e-16
    w = ex / den.index_select(0, dst)
    msg = (v.index_select(0, src) + e) * w.unsqueeze(-1)
    return scatter_sum(msg, dst, n_dst)


def gt_attention_online(q, k, v, e, row, colptr) -> Tensor:
    """triton/gt.py:81-179 — per-dst online softmax over CSC edges, fp32; zero in-degree ⇒ zeros.

    Pure-Python loop: small cases only (independent cross-check of ``gt_attention``).
    """
    n_dst, H, C = q.shape
    out = torch.zeros((n_dst, H, C), dtype=torch.float32)
    scale = 1.0 / math.sqrt(float(C))
    for d in range(n_dst):
        s, t = int(colptr[d]), int(colptr[d + 1])
        if s == t:
            continue
        acc = torch.zeros((H, C))
        l_i = torch.zeros(H)
        m_i = torch.full((H,), -float("inf"))
        for ei in range(s, t):
            j = int(row[ei])
            ke = k[j].float() + e[ei].float()
            ve = v[j].float() + e[ei].float()
            qk = (q[d].float() * ke).sum(-1) * scale
            m_ij = torch.maximum(m_i, qk)
            a = torch.exp(qk - m_ij)
            corr = torch.exp(m_i - m_ij)
            acc = acc * corr[:, None] + a[:, None] * ve
            l_i = l_i * corr + a
            m_i = m_ij
        out[d] = acc / l_i[:, None]
    return out


def _gt_core(sd, prefix, xs_n, xd_n, x_dst_skip, edge_attr, edge_index, num_heads, cond_dst=None):
    """layers/block.py:623-687 (get_qkve, heads reshape, conv) + projection/residual/MLP tail
    shared by :1019-1029 (mapper) and :1268-1271 (processor)."""
    n_dst = xd_n.shape[0]
    x_r = _linear(sd, prefix + ".lin_self", xd_n)
    q = _linear(sd, prefix + ".lin_query", xd_n)
    k = _linear(sd, prefix + ".lin_key", xs_n)
    v = _linear(sd, prefix + ".lin_value", xs_n)
    if prefix + ".edge_pre_mlp.0.weight" in sd:  # block.py:575-583 — Linear + activation
        edge_attr = F.gelu(_linear(sd, prefix + ".edge_pre_mlp.0", edge_attr))
    e = _linear(sd, prefix + ".lin_edge", edge_attr)
    H = num_heads
    q, k, v, e = (t.reshape(t.shape[0], H, t.shape[1] // H) for t in (q, k, v, e))  # explicit width: an empty edge list stays reshapeable
    if prefix + ".q_norm.weight" in sd:  # block.py:655-660 — AutocastLayerNorm(Ch, bias=False)
        q = F.layer_norm(q, (q.shape[-1],), sd[prefix + ".q_norm.weight"], None, 1e-5).type_as(q)
        k = F.layer_norm(k, (k.shape[-1],), sd[prefix + ".k_norm.weight"], None, 1e-5).type_as(k)
    att = gt_attention(q, k, v, e, edge_index, n_dst).reshape(n_dst, -1)
    out = _linear(sd, prefix + ".projection", att + x_r) + x_dst_skip
    h = _layer_norm(sd, prefix + ".layer_norm_mlp_dst", out, cond=cond_dst)
    return mlp(sd, prefix + ".node_dst_mlp", h) + out


def gt_processor_block(sd, prefix, x, edge_attr, edge_index, num_heads, cond=None):
    """layers/block.py:1219-1273 — returns nodes_new (edge_attr is returned unchanged by the reference); ``cond`` feeds both LayerNorms
    when they are ConditionalLayerNorm kernels (:1233-1271)."""
    xn = _layer_norm(sd, prefix + ".layer_norm_attention", x, cond=cond)
    return _gt_core(sd, prefix, xn, xn, x, edge_attr, edge_index, num_heads, cond_dst=cond)


def gt_mapper_block(sd, prefix, x_src, x_dst, edge_attr, edge_index, num_heads, cond=None):
    """layers/block.py:963-1029 — separate LayerNorm for src (:940) and dst; update_src_nodes=False; ``cond`` = (cond_src, cond_dst) (:978-980)."""
    cs, cd = cond if cond is not None else (None, None)
    xs_n = _layer_norm(sd, prefix + ".layer_norm_attention_src", x_src, cond=cs)
    xd_n = _layer_norm(sd, prefix + ".layer_norm_attention", x_dst, cond=cd)
    return _gt_core(sd, prefix, xs_n, xd_n, x_dst, edge_attr, edge_index, num_heads, cond_dst=cd)


def gt_processor(sd, x, edge_attr, edge_index, num_layers, num_heads, max_layers=None, cond=None):
    """layers/processor.py:552-626 — every layer re-projects the raw edge_attr with its own lin_edge."""
    for layer in range(num_layers if max_layers is None else min(num_layers, max_layers)):
        x = gt_processor_block(sd, f"proc.{layer}", x, edge_attr, edge_index, num_heads, cond=cond)
    return x


def gt_forward_mapper(sd, x_src, x_dst, edge_attr, edge_index, num_heads, cond=None):
    """layers/mapper.py:335-386, 480-597 — returns (x_src unchanged, x_dst); chunking does not change values."""
    xs = _linear(sd, "emb_nodes_src", x_src)
    xd = _linear(sd, "emb_nodes_dst", x_dst)
    return x_src, gt_mapper_block(sd, "proc", xs, xd, edge_attr, edge_index, num_heads, cond=cond)


def gt_backward_mapper(sd, x_src, x_dst, edge_attr, edge_index, num_heads):
    """layers/mapper.py:600-704 — dst embedded by emb_nodes_dst, extractor = LayerNorm + Linear (:688-690)."""
    xd = _linear(sd, "emb_nodes_dst", x_dst)
    out = gt_mapper_block(sd, "proc", x_src, xd, edge_attr, edge_index, num_heads)
    h = F.layer_norm(out, (out.shape[-1],), sd["node_data_extractor.0.weight"], sd["node_data_extractor.0.bias"], 1e-5)
    return _linear(sd, "node_data_extractor.1", h)


# ----------------------------------------------------------------------------------------------
# whole encoder → processor → decoder step (the bench "step"), models/encoder_processor_decoder.py:260-324
# ----------------------------------------------------------------------------------------------
def gt_encode_process_decode(sds, graph, x_grid, x_mesh, num_layers, num_heads, max_layers=None):
    """encoder (data→hidden), processor, latent skip (:295-296), decoder (hidden→data)."""
    _, lat = gt_forward_mapper(sds["encoder"], x_grid, x_mesh, graph["enc_attr"], graph["enc_index"], num_heads)
    proc = gt_processor(sds["processor"], lat, graph["proc_attr"], graph["proc_index"], num_layers, num_heads, max_layers)
    proc = proc + lat
    return gt_backward_mapper(sds["decoder"], proc, x_grid, graph["dec_attr"], graph["dec_index"], num_heads)


def gnn_encode_process_decode(sds, graph, x_grid, x_mesh, num_layers, autocast_ln=False, max_layers=None):
    """GNN variant: the decoder's dst input is the encoder's updated src embedding (:260-269, :316-318)."""
    src_emb, lat = gnn_forward_mapper(sds["encoder"], x_grid, x_mesh, graph["enc_attr"], graph["enc_index"], autocast_ln)
    proc = gnn_processor(sds["processor"], lat, graph["proc_attr"], graph["proc_index"], num_layers, autocast_ln, max_layers)
    proc = proc + lat
    return gnn_backward_mapper(sds["decoder"], proc, src_emb, graph["dec_attr"], graph["dec_index"], autocast_ln)


# ----------------------------------------------------------------------------------------------
# graph providers, node attributes and the model glue either side of the step
# ----------------------------------------------------------------------------------------------
def static_graph_provider_edges(edge_index: Tensor, attrs: list[Tensor], trainable: Optional[Tensor], src_size: int, dst_size: int,
                                batch_size: int) -> tuple[Tensor, Tensor]:  # fmt: skip
    """layers/graph_provider.py:185-256 + layers/graph.py:38-46 — stable dst sort once, cat(fixed attributes)[perm], cat with the trainable
    tensor, rows tiled ``batch_size`` times; edge_index replicated with (src_size, dst_size) offsets per batch element."""
    ei, perm = sort_edge_index_by_dst(edge_index)
    ea = torch.cat(attrs, dim=1).index_select(0, perm)
    if trainable is not None:
        ea = torch.cat([ea, trainable], dim=-1)
    ea = ea.repeat(batch_size, 1)
    inc = torch.tensor([[src_size], [dst_size]], dtype=torch.int64)
    return ea, torch.cat([ei + i * inc for i in range(batch_size)], dim=1)


def named_node_attributes(coords: Tensor, trainable: Optional[Tensor], batch_size: int) -> Tensor:
    """layers/graph.py:96-118 — [sin(coords), cos(coords), trainable], tiled over the batch."""
    a = torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)
    if trainable is not None:
        a = torch.cat([a, trainable], dim=-1)
    return a.repeat(batch_size, 1)


def assemble_input(x: Tensor, node_attrs: Tensor) -> Tensor:
    """models/encoder_processor_decoder.py:115-125 — "batch time ensemble grid vars -> (batch ensemble grid) (time vars)" ++ attributes."""
    b, t, e, g, v = x.shape
    return torch.cat([x.permute(0, 2, 3, 1, 4).reshape(b * e * g, t * v), node_attrs], dim=-1)


def assemble_output(x_out: Tensor, x: Tensor, batch: int, ensemble: int, n_step_output: int, in_prog: list[int], out_prog: list[int],
                    relu_vars: list[int], step: int = -1) -> Tensor:  # fmt: skip
    """models/encoder_processor_decoder.py:129-163 — back to (batch time ensemble grid vars), SkipConnection residual (layers/residual.py:60-81:
    the input's ``step`` slice for every output step) on the prognostic variables, then ReluBounding (layers/bounding.py:81-86)."""
    g = x_out.shape[0] // (batch * ensemble)
    y = x_out.reshape(batch, ensemble, g, n_step_output, -1).permute(0, 3, 1, 2, 4).to(x.dtype).clone()
    skip = x[:, step].unsqueeze(1).expand(-1, n_step_output, -1, -1, -1)
    y[..., out_prog] += skip[..., in_prog]
    if relu_vars:
        y[..., relu_vars] = F.relu(y[..., relu_vars])
    return y


def anemoi_model_forward(kind: str, sd: StateDict, fx: dict, num_layers: int = 2) -> Tensor:
    """``AnemoiModelEncProcDec.forward`` (models/encoder_processor_decoder.py:185-330), single dataset "data", no model sharding, driven by the
    reference model ``state_dict`` and the fixture's raw graph (``tests/golden/model_forward.pt``)."""
    d, x = fx["dims"], fx["x"]
    batch = x.shape[0]

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}

    def edges(name, provider_prefix, n_src, n_dst):
        gr = fx["graph"][name]
        return static_graph_provider_edges(gr["edge_index"], [gr["edge_length"], gr["edge_dirs"]], sd.get(provider_prefix + "trainable.trainable"),
                                           n_src, n_dst, batch)  # fmt: skip

    attrs_data = named_node_attributes(fx["coords"]["data"], sd.get("node_attributes.trainable_tensors.data.trainable"), batch)
    attrs_hid = named_node_attributes(fx["coords"]["hidden"], sd.get("node_attributes.trainable_tensors.hidden.trainable"), batch)
    x_data = assemble_input(x, attrs_data)
    enc_attr, enc_index = edges("enc", "encoder_graph_provider.data.", d["n_data"], d["n_hid"])
    proc_attr, proc_index = edges("proc", "processor_graph_provider.", d["n_hid"], d["n_hid"])
    dec_attr, dec_index = edges("dec", "decoder_graph_provider.data.", d["n_hid"], d["n_data"])
    enc, proc, dec = sub("encoder.data."), sub("processor."), sub("decoder.data.")
    if kind == "graphtransformer":
        x_data_latent, lat = gt_forward_mapper(enc, x_data, attrs_hid, enc_attr, enc_index, d["heads"])
        p = gt_processor(proc, lat, proc_attr, proc_index, num_layers, d["heads"]) + lat
        out = gt_backward_mapper(dec, p, x_data_latent, dec_attr, dec_index, d["heads"])
    else:
        x_data_latent, lat = gnn_forward_mapper(enc, x_data, attrs_hid, enc_attr, enc_index)
        p = gnn_processor(proc, lat, proc_attr, proc_index, num_layers) + lat
        out = gnn_backward_mapper(dec, p, x_data_latent, dec_attr, dec_index)
    return assemble_output(out, x, batch, x.shape[2], d["t_out"], fx["in_prog"], fx["out_prog"], fx["bound_vars"])
