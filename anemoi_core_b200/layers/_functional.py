"""Shared forward plumbing: compute-dtype choice, packed (cast / concatenated / K-padded) weights, fused MLP runs,
and the per-graph CSR cache.  Everything numerical is a call into ``ops`` (the C ABI)."""

from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable
from typing import Optional
from typing import Sequence

import torch
from torch import Tensor
from torch import nn

from .. import ops
from .._ident import version

SUPPORTED = (torch.float32, torch.bfloat16)


# fp16 autocast (the reference's default training precision, "16-mixed": training/config/training/single.yaml:30): there are no fp16 kernels here.
# "bf16" (default): compute as under bf16 autocast, with a one-time warning - same tensor-core path, fp32 accumulation, fp32's exponent range
# (a GradScaler's loss scale is harmless), 8 instead of 11 significand bits in the stored activations; "fp32": the fp32 path (exact bf16 x 3 split
# on the tensor cores: at least fp16's accuracy, ~5x slower than bf16); "error": refuse.
FP16_AUTOCAST = os.environ.get("ANEMOI_B200_FP16_AUTOCAST", "bf16")
_warned_fp16 = False


def compute_dtype(*tensors: Tensor) -> torch.dtype:
    """bf16 under ``torch.autocast(dtype=bfloat16)`` or for bf16 inputs (tcgen05 path); fp32 otherwise (parity mode); fp16 autocast: FP16_AUTOCAST."""
    global _warned_fp16
    if any(t.dtype == torch.float16 for t in tensors):
        raise NotImplementedError("float16 tensors are not supported by the B200 kernels (float32 and bfloat16 are): cast them with .to(torch.bfloat16); "
                                  "under fp16 autocast the modules compute in bfloat16 and return bfloat16 (ANEMOI_B200_FP16_AUTOCAST)")  # fmt: skip
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        if dt == torch.float16 and FP16_AUTOCAST in ("bf16", "fp32"):
            if not _warned_fp16:
                import warnings

                warnings.warn(f"anemoi_core_b200: fp16 autocast requested; the B200 path has no fp16 kernels and computes in "
                              f"{'bfloat16' if FP16_AUTOCAST == 'bf16' else 'float32'} instead (ANEMOI_B200_FP16_AUTOCAST=bf16|fp32|error)", stacklevel=3)  # fmt: skip
                _warned_fp16 = True
            dt = torch.bfloat16 if FP16_AUTOCAST == "bf16" else torch.float32
    else:
        dt = tensors[0].dtype
        for t in tensors[1:]:
            if t.dtype != dt:
                dt = torch.promote_types(dt, t.dtype)
    if dt not in SUPPORTED:
        raise NotImplementedError(f"compute dtype {dt} is not implemented (float32 and bfloat16 are)")
    return dt


def round8(k: int) -> int:
    return (k + 7) // 8 * 8


def pad_k(k: int, dt: torch.dtype) -> int:
    """K of a bf16 GEMM operand: multiple of 8 (16-byte TMA row stride) and at least 64 (one 128-byte swizzle span), so that
    even the edge_dim = 3..11 and in_channels = 12 embeddings run on the tcgen05 kernel instead of the FFMA one; fp32 untouched."""
    if dt != torch.bfloat16:
        return k
    return max(64, round8(k))


class WeightPack:
    """Derived tensors of a module's parameters (dtype casts, row-concatenations, zero K-padding), rebuilt when a
    source parameter changes (in-place update, ``load_state_dict``, ``.to()``)."""

    def __init__(self) -> None:
        self._store: dict = {}
        self.frozen = False  # inference: skip the per-call parameter-version check (see freeze_packed_weights)

    def get(self, key, sources: Sequence[Optional[Tensor]], build: Callable[[], Tensor]) -> Tensor:
        if self.frozen:
            hit = self._store.get(key)
            if hit is not None:
                return hit[1]
        sig = tuple(None if t is None else (t.data_ptr(), version(t)) for t in sources)  # storage identity + in-place version counter
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad(), torch.autocast("cuda", enabled=False):  # derived weights are built in fp32, never under autocast
            val = build()
        self._store[key] = (sig, val)
        return val

    def weight(self, layers: Sequence[nn.Module], dt: torch.dtype, cols: Optional[slice] = None) -> Tensor:
        """Row-concatenated weights of ``layers`` ([sum N_i, K]) in dtype ``dt``; for bf16 the K dimension is zero-padded
        to a multiple of 8 (TMA needs 16-byte row strides).  ``cols`` selects input columns (split of a concatenated input)."""
        ws = [l.weight for l in layers]

        def build() -> Tensor:
            w = torch.cat([x.detach() for x in ws], 0) if len(ws) > 1 else ws[0].detach()
            if cols is not None:
                w = w[:, cols]
            k = w.shape[1]
            if pad_k(k, dt) != k:
                w = torch.nn.functional.pad(w, (0, pad_k(k, dt) - k))
            return w.to(dt).contiguous()

        return self.get(("w", tuple(id(l) for l in layers), dt, None if cols is None else (cols.start, cols.stop)), ws, build)

    def bias(self, layers: Sequence[nn.Module]) -> Optional[Tensor]:
        bs = [getattr(l, "bias", None) for l in layers]
        if all(b is None for b in bs):
            return None

        def build() -> Tensor:
            parts = [b.detach().float() if b is not None else torch.zeros(l.weight.shape[0], device=l.weight.device) for b, l in zip(bs, layers)]
            return torch.cat(parts).contiguous()

        return self.get(("b", tuple(id(l) for l in layers)), bs, build)

    def f32(self, p: Optional[Tensor]) -> Optional[Tensor]:
        if p is None:
            return None
        if p.dtype == torch.float32 and p.is_contiguous():
            return p.detach()
        return self.get(("f32", id(p)), [p], lambda: p.detach().float().contiguous())


class PackOwner(nn.Module):
    """Base of the modules that own a ``WeightPack`` (``self._pack``).  Derived weights are revalidated against (storage pointer, version counter)
    of their source parameters on every call — but an in-place edit through ``param.data`` (some EMA / weight-swapping utilities) moves neither.
    As a second line of defence every switch between training and evaluation mode drops the derived tensors, so that an evaluation phase always
    starts from the current parameters; a FROZEN pack (``freeze_packed_weights``: the user promised not to mutate, and a captured CUDA graph may
    hold the derived tensors' addresses) is left alone.  ``invalidate_packed_weights`` does the same on demand."""

    def train(self, mode: bool = True):
        pack = getattr(self, "_pack", None)
        if isinstance(pack, WeightPack) and mode != self.training and not pack.frozen:
            pack._store.clear()
        return super().train(mode)


def invalidate_packed_weights(module: nn.Module) -> nn.Module:
    """Drop every derived (cast / concatenated / folded) weight under ``module``; the next forward rebuilds them from the parameters.  For code
    that edits parameters through ``.data`` (which bumps no version counter).  Not while a captured CUDA graph of the module is alive: the graph
    reads the derived tensors at their captured addresses."""
    for m in module.modules():
        pack = getattr(m, "_pack", None)
        if isinstance(pack, WeightPack):
            pack._store.clear()
    return module


def freeze_packed_weights(module: nn.Module, frozen: bool = True) -> nn.Module:
    """Inference switch: stop re-validating the derived (cast / concatenated / folded) weights against the parameters on every
    call (a tuple of data_ptr / _version per source tensor, a few microseconds per GEMM on the host).  Call again with False —
    or after it, mutate nothing — before changing parameters."""
    for m in module.modules():
        pack = getattr(m, "_pack", None)
        if isinstance(pack, WeightPack):
            pack.frozen = frozen
    return module


def as_operand(x: Tensor, dt: torch.dtype, k_weight: int) -> Tensor:
    """Make ``x`` [M, K] a GEMM A-operand of dtype ``dt`` whose width matches the (possibly K-padded) weight."""
    if x.dim() != 2:
        raise ValueError(f"expected [nodes, channels], got {tuple(x.shape)}")
    k = x.shape[1]
    if x.dtype == dt and k == k_weight and x.stride(1) == 1:
        return x
    return ops.cast_pad(x, dt, k_weight)


def fused_linear(pack: WeightPack, x: Tensor, layers: Sequence[nn.Module], dt: torch.dtype, cols: Optional[slice] = None, **kw) -> Tensor:
    """``x @ cat(W_i).T + cat(b_i)`` with the epilogue options of ``ops.linear`` (gelu / residual / gathers / out)."""
    for l in layers:
        if not hasattr(l, "weight") or l.weight.dim() != 2:
            raise NotImplementedError(f"{type(l).__name__} is not a Linear-like parameter container (needs a 2-D .weight)")
    w = pack.weight(layers, dt, cols)
    bias = pack.bias(layers) if kw.pop("use_bias", True) else None
    return linear_with_stats(as_operand(x, dt, w.shape[1]), w, bias, **kw)


def linear_with_stats(a: Tensor, w: Tensor, bias: Optional[Tensor], want_stats: bool = False, **kw) -> Tensor:
    """``ops.linear``; with ``want_stats`` (and an output a LayerNorm can be folded over) the GEMM epilogue also writes the row statistics
    of its output and the result is tagged with them (``tag_row_stats``) for the GEMM that will normalise it."""
    out_dt = kw["out"].dtype if kw.get("out") is not None else (kw.get("out_dtype") or a.dtype)
    if not (want_stats and FUSED_ROW_STATS and wants_row_stats(w.shape[0], out_dt)):
        return ops.linear(a, w, bias, **kw)
    stats = ops.partial_stats_buffer(a.shape[0], w.shape[0], a.device)
    return tag_row_stats(ops.linear(a, w, bias, stats_out=stats, **kw), stats)


# ---- row statistics handed from the GEMM that produces a tensor to the GEMM that normalises it ------------------------------------
# The producer's epilogue writes per-row partial (sum, sum of squares) next to its output (``ops.linear(stats_out=)``); the pair travels as
# an attribute of the output tensor object and is honoured only while that tensor is unchanged (same storage, version, shape), so a slice,
# a gather or an in-place update silently falls back to the ``row_stats`` pass.
# A/B switch.  Default OFF since round 2: measured on the same box (profiles/r2/call8_ab_row_stats.txt) the separate row_stats pass gives
# 8.64 ms / step against 8.89-8.94 ms with the statistics written by the producing GEMM's epilogue - the epilogue is the GEMMs' critical
# path, and every instruction added to it (shifted sums, the per-block division) costs more than the 16 us pass it replaces.
FUSED_ROW_STATS = os.environ.get("ANEMOI_B200_FUSED_ROW_STATS", "0") != "0"


def tag_row_stats(t: Tensor, stats: Tensor) -> Tensor:
    t._anemoi_row_stats = (stats, t.data_ptr(), version(t), tuple(t.shape), tuple(t.stride()))
    return t


def tagged_row_stats(t: Tensor) -> Optional[Tensor]:
    tag = getattr(t, "_anemoi_row_stats", None)
    if tag is None or tag[1:] != (t.data_ptr(), version(t), tuple(t.shape), tuple(t.stride())):
        return None
    return tag[0]


def wants_row_stats(n_out: int, dt: torch.dtype) -> bool:
    """Can a LayerNorm over ``n_out`` columns of a ``dt`` GEMM output be folded into the next GEMM (same test as ``can_fold_ln``)?"""
    return dt == torch.bfloat16 and n_out % 8 == 0 and 64 <= n_out <= 2048


def can_fold_ln(ln: nn.Module, k: int, dt: torch.dtype) -> bool:
    """A LayerNorm directly in front of a tcgen05 GEMM is folded into it (bf16 path only; the fp32 parity path keeps the kernel)."""
    return dt == torch.bfloat16 and isinstance(ln, torch.nn.LayerNorm) and k % 8 == 0 and 64 <= k <= 2048


def ln_linear(pack: WeightPack, x: Tensor, ln: nn.Module, key, sources: Sequence[Optional[Tensor]], build32: Callable[[], tuple], dt: torch.dtype,
              cond: Optional[Tensor] = None, **kw) -> Tensor:  # fmt: skip
    """``linear(LayerNorm(x))`` for the weight / bias produced by ``build32() -> (w fp32 [N, K], b fp32 [N] | None)``.

    bf16: the LayerNorm is folded into the GEMM — W' = W * gamma, bias' = b + W beta, colsum = sum_k bf16(W') — and only the per-row
    (mean, rstd) are computed beforehand (``ops.row_stats`` on the bf16 operand, so constant rows cancel exactly); the normalised
    activations are never written.  fp32 (parity mode) or unsupported shapes: LayerNorm kernel, then the plain GEMM."""
    from .normalization import ConditionalLayerNorm
    from .normalization import _check_plain_layernorm

    k = x.shape[1]
    if isinstance(ln, ConditionalLayerNorm):
        # per-row affine from the conditioning: cannot be folded into the weights; one conditional-LayerNorm kernel, then the plain GEMM
        def build_cond():
            w32, b32 = build32()
            kk = w32.shape[1]
            if pad_k(kk, dt) != kk:
                w32 = torch.nn.functional.pad(w32, (0, pad_k(kk, dt) - kk))
            return w32.to(dt).contiguous(), (None if b32 is None else b32.contiguous())

        w, b = pack.get(("ln_cond", key, dt), list(sources), build_cond)
        return linear_with_stats(as_operand(ln.run(x, cond, dt), dt, w.shape[1]), w, b, **kw)
    _check_plain_layernorm(ln)
    srcs = list(sources) + [ln.weight, ln.bias]
    if can_fold_ln(ln, k, dt):

        def build():
            w32, b32 = build32()
            gamma = ln.weight.detach().float() if ln.weight is not None else torch.ones(k, device=w32.device)
            beta = ln.bias.detach().float() if ln.bias is not None else torch.zeros(k, device=w32.device)
            wf = (w32 * gamma).to(dt).contiguous()
            bias = (b32 if b32 is not None else torch.zeros(w32.shape[0], device=w32.device)) + w32 @ beta
            return wf, bias.contiguous(), wf.float().sum(1).contiguous()

        wf, bias, colsum = pack.get(("ln_fold", key, dt), srcs, build)
        a = as_operand(x, dt, k)
        st = tagged_row_stats(a)
        if st is not None:  # produced by the epilogue of the GEMM that wrote ``a``: no statistics pass at all
            return linear_with_stats(a, wf, bias, ln_stats=st, ln_dim=k, ln_eps=ln.eps, ln_colsum=colsum, **kw)
        return linear_with_stats(a, wf, bias, ln_stats=ops.row_stats(a, ln.eps), ln_colsum=colsum, **kw)

    def build_plain():
        w32, b32 = build32()
        kk = w32.shape[1]
        if pad_k(kk, dt) != kk:
            w32 = torch.nn.functional.pad(w32, (0, pad_k(kk, dt) - kk))
        return w32.to(dt).contiguous(), (None if b32 is None else b32.contiguous())

    w, b = pack.get(("ln_plain", key, dt), srcs, build_plain)
    xn = ops.layer_norm(x, pack.f32(ln.weight), pack.f32(ln.bias), ln.eps, out_dtype=dt)
    return linear_with_stats(as_operand(xn, dt, w.shape[1]), w, b, **kw)


def cat_linear32(layers: Sequence[nn.Module]) -> tuple:
    """fp32 (weight, bias) of row-concatenated Linear containers."""
    w = torch.cat([l.weight.detach().float() for l in layers], 0)
    if all(getattr(l, "bias", None) is None for l in layers):
        return w, None
    b = torch.cat([l.bias.detach().float() if l.bias is not None else torch.zeros(l.weight.shape[0], device=w.device) for l in layers])
    return w, b


def linear_sources(layers: Sequence[nn.Module]) -> list:
    return [l.weight for l in layers] + [getattr(l, "bias", None) for l in layers]


def layer_norm_mod(pack: WeightPack, ln: nn.Module, x: Tensor, dt: torch.dtype, residual: Optional[Tensor] = None, groups: int = 1) -> Tensor:
    from .normalization import _check_plain_layernorm

    _check_plain_layernorm(ln)
    return ops.layer_norm(x, pack.f32(ln.weight), pack.f32(ln.bias), ln.eps, residual=residual, out_dtype=dt, groups=groups)


# ------------------------------------------------------------------------------------------------------------
# per-graph CSR cache
# ------------------------------------------------------------------------------------------------------------
_CSR_CACHE: "OrderedDict[tuple, tuple[Tensor, ops.GraphCSR]]" = OrderedDict()
_CSR_CACHE_SIZE = 16


def csr_for(edge_index: Tensor, n_src: int, n_dst: int) -> ops.GraphCSR:
    """CSR plan of a dst-sorted edge_index, cached on the tensor's identity (storage pointer, version, shape).  The
    cache holds a reference to the tensor so its storage cannot be recycled under the key.  Built once per static
    graph; the reference rebuilds its CSC (two sorts) per layer per forward (layers/block.py:779-782)."""
    key = (edge_index.data_ptr(), version(edge_index), tuple(edge_index.shape), tuple(edge_index.stride()), str(edge_index.device), n_src, n_dst)
    hit = _CSR_CACHE.get(key)
    if hit is not None:
        _CSR_CACHE.move_to_end(key)
        return hit[1]
    validate = not torch.cuda.is_current_stream_capturing()
    csr = ops.build_csr(edge_index, n_src, n_dst, validate=validate)
    _CSR_CACHE[key] = (edge_index, csr)
    while len(_CSR_CACHE) > _CSR_CACHE_SIZE:
        _CSR_CACHE.popitem(last=False)
    return csr


def pad_edge_attr(edge_attr: Tensor, width: int = 0) -> Tensor:
    """fp32 [E, width] zero-padded copy of the raw edge attributes (``width`` = 16 for the folded attention path: one
    64-byte row = two full sectors per edge; default ceil4(d_e))."""
    d = edge_attr.shape[1]
    w = width or (d + 3) // 4 * 4
    if edge_attr.dtype == torch.float32 and d == w and edge_attr.is_contiguous():
        return edge_attr
    return ops.cast_pad(edge_attr, torch.float32, w)
