"""Tensor-level wrappers over the C ABI (``include/anemoi_b200.h``): each function checks shapes, passes raw device
pointers + leading dimensions + the current CUDA stream, and returns torch tensors.  No compute happens in Python
or PyTorch here; a CPU tensor is an error (there is no fallback path).
"""

from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import BF16
from ._lib import EPI_GELU
from ._lib import EPI_REVERSE
from ._lib import F32

__all__ = ["GraphCSR", "build_csr", "layer_norm", "row_stats", "linear", "gt_attention", "graphconv_ln_aggregate", "graphconv_fused", "cast_pad", "add", "dtype_code"]


# ---- instrumentation: launch counter and optional CUDA-event timing of every C-ABI call (bench.py roofline leg) -------
LAUNCHES = 0  # kernels launched through the C ABI since import (each entry point launches exactly one kernel)
_TIMER = None  # None or a list collecting (name, start_event, end_event, flops, bytes)


def start_timing() -> list:
    global _TIMER
    _TIMER = []
    return _TIMER


def stop_timing() -> list:
    global _TIMER
    rec, _TIMER = _TIMER, None
    return rec or []


# ---- traversal hint (L2 scheduling; results never depend on it) ---------------------------------------------------------------
# A streaming kernel leaves the rows it touched LAST in L2 (126 MB), the next kernel usually starts at row 0, which was evicted long ago.
# ``set_traversal`` lets a caller that knows the producer / consumer chain (the GraphTransformer block, layers/block.py) make the next
# GEMM / attention (``gemm``) or row-statistics pass (``stats``) walk the rows from the bottom up, so that it starts where its
# predecessor ended: the "serpentine" schedule.  Thread-local, consumed by ``linear`` / ``gt_attention`` / ``row_stats``.
_TRAVERSAL = threading.local()


def set_traversal(gemm: bool = False, stats: bool = False) -> None:
    _TRAVERSAL.gemm, _TRAVERSAL.stats = bool(gemm), bool(stats)


def _rev(kind: str) -> int:
    return EPI_REVERSE if getattr(_TRAVERSAL, kind, False) else 0


class _Timed:
    """Scope of ONE C-ABI call: counts the launch, optionally brackets it with CUDA events, and makes the device of the call's tensors
    (recorded by ``_need_cuda``) the current device for its duration — the library launches on the calling thread's current device and
    ``_stream()`` hands it that device's current stream, so a model on cuda:1 works without ``torch.cuda.set_device`` like any torch op."""

    __slots__ = ("name", "flops", "bytes", "ev", "guard")

    def __init__(self, name: str, flops: float = 0.0, nbytes: float = 0.0):
        self.name, self.flops, self.bytes, self.ev, self.guard = name, flops, nbytes, None, None

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += 1
        dev = getattr(_CALL, "dev", None)
        if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
            self.guard = torch.cuda.device(dev)
            self.guard.__enter__()
        if _TIMER is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _TIMER.append((self.name, self.ev, end, self.flops, self.bytes))
        if self.guard is not None:
            self.guard.__exit__(*exc)
        return False


_CALL = threading.local()  # device of the tensors of the C-ABI call being prepared on this thread (set by _need_cuda)


def _nbytes(*ts: Optional[Tensor]) -> float:
    return float(sum(t.shape[0] * t.shape[1] * t.element_size() if t.dim() == 2 else t.numel() * t.element_size() for t in ts if t is not None))


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"anemoi_core_b200 computes in float32 or bfloat16, got {dt}")


def _need_cuda(*tensors: Optional[Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("anemoi_core_b200 kernels run on CUDA tensors only (sm_100a); got a CPU tensor and there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is not None:
        _CALL.dev = dev
    return dev


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: Tensor) -> tuple[int, int, int]:
    """(rows, cols, leading dimension) of a 2-D tensor whose last dimension is contiguous."""
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError("last dimension must be contiguous")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t.shape[0], t.shape[1], ld


def _f32(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("bias / LayerNorm / edge parameters must be contiguous float32 tensors")
    return t


# ------------------------------------------------------------------------------------------------------------
# integer path
# ------------------------------------------------------------------------------------------------------------
@dataclass
class GraphCSR:
    """Cached per-graph index plan (replaces the per-layer ``edge_index_to_csc`` of layers/block.py:779-782)."""

    n_src: int
    n_dst: int
    n_edges: int
    colptr: Tensor  # int64 [n_dst+1]  — reference dtype (triton/utils.py:61)
    colptr32: Tensor  # int32 [n_dst+1]
    src32: Tensor  # int32 [E]
    dst32: Tensor  # int32 [E]
    tiles: object = None  # AttnTilePlan, False (no plan exists) or None (not built yet); see attention_tiles()
    rev: object = None  # (rev_ptr32 [n_src + 1], rev_eid32 [E]): edges stably sorted by source, for the backward's src-major pass


@dataclass
class AttnTilePlan:
    """Destination-tile plan of the tiled attention kernel (include/anemoi_b200.h: anemoi_b200_attn_tile_plan)."""

    n_tiles: int
    n_slots: int
    tile_meta: Tensor  # int32 [n_tiles, 4] device
    slot_src: Tensor  # int32 [n_slots] device
    emeta: Tensor  # int16 [E] device: slot | row << 8 of every edge inside its tile
    max_edges: int  # edges-per-tile bound the plan was built with

    @property
    def reuse(self) -> float:
        """edges per gathered source row: how often a tile re-uses a row it fetched (1.0 = no locality)."""
        return float(self.emeta.numel()) / max(self.n_slots, 1)


ATTN_TILE_MAX_EDGES = 168  # fits every dp <= 16: min(224, 10752 // (4 * 16))


def plan_attention_tiles_host(src32, colptr32, n_src: int, n_dst: int, max_edges: int = ATTN_TILE_MAX_EDGES):
    """Run the host-side planner on numpy int32 arrays; returns (tile_meta [T,4], slot_src [S], emeta [E] uint16) numpy arrays or None."""
    import ctypes

    import numpy as np

    src32 = np.ascontiguousarray(src32, dtype=np.int32)
    colptr32 = np.ascontiguousarray(colptr32, dtype=np.int32)
    n_edges = int(colptr32[n_dst]) if colptr32.size else 0
    tile_meta = np.zeros((max(n_dst, 1), 4), dtype=np.int32)
    slot_src = np.zeros(max(n_edges, 1), dtype=np.int32)
    emeta = np.zeros(max(n_edges, 1), dtype=np.uint16)
    nt, ns = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = _lib.load().anemoi_b200_attn_tile_plan(src32.ctypes.data, colptr32.ctypes.data, n_src, n_dst, max_edges, tile_meta.ctypes.data,
                                               slot_src.ctypes.data, emeta.ctypes.data, ctypes.addressof(nt), ctypes.addressof(ns))  # fmt: skip
    if rc == -3:
        return None
    _lib.check(rc, "anemoi_b200_attn_tile_plan")
    return tile_meta[: nt.value].copy(), slot_src[: ns.value].copy(), emeta[:n_edges].copy()


def attention_tiles(csr: GraphCSR) -> Optional[AttnTilePlan]:
    """Tile plan of ``csr`` (built once on the host, cached on the CSR object); None when some destination has too many sources."""
    if csr.tiles is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # needs a device -> host copy: build it in the warm-up pass, not under capture
        res = plan_attention_tiles_host(csr.src32.cpu().numpy(), csr.colptr32.cpu().numpy(), csr.n_src, csr.n_dst)
        if res is None:
            csr.tiles = False
        else:
            dev = csr.colptr32.device
            tm, ss = (torch.from_numpy(a).to(dev) for a in res[:2])
            em = torch.from_numpy(res[2].view("int16")).to(dev)
            csr.tiles = AttnTilePlan(int(tm.shape[0]), int(ss.numel()), tm.contiguous(), ss, em, ATTN_TILE_MAX_EDGES)
    return csr.tiles or None


def build_csr(edge_index: Tensor, n_src: int, n_dst: int, validate: bool = True) -> GraphCSR:
    """CSR of a dst-sorted ``edge_index`` [2, E] int64.  ``validate`` synchronises once to raise on unsorted /
    out-of-range input (the reference trusts ``edges_are_dst_sorted``; we check once per cached graph)."""
    _need_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise TypeError("edge_index must be an int64 tensor of shape [2, E]")
    edge_index = edge_index.contiguous()
    n_edges = edge_index.shape[1]
    dev = edge_index.device
    colptr = torch.empty(n_dst + 1, dtype=torch.int64, device=dev)
    colptr32 = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    src32 = torch.empty(n_edges, dtype=torch.int32, device=dev)
    dst32 = torch.empty(n_edges, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    with _Timed("csr_build", 0.0, 24.0 * n_edges + 12.0 * n_dst):
        rc = lib.anemoi_b200_csr_build(
            _ptr(edge_index), n_edges, n_src, n_dst, _ptr(colptr), _ptr(colptr32), _ptr(src32), _ptr(dst32), _ptr(status), _stream()
        )
    _lib.check(rc, "anemoi_b200_csr_build")
    if validate:
        st = int(status.item())
        if st & 2:
            raise ValueError(f"edge_index has node ids outside [0, {n_src}) x [0, {n_dst})")
        if st & 1:
            raise ValueError("edge_index is not sorted by destination; sort it (sort_edge_index_by_dst) or pass edges_are_dst_sorted=False")
    return GraphCSR(n_src, n_dst, n_edges, colptr, colptr32, src32, dst32)


# ------------------------------------------------------------------------------------------------------------
# float kernels
# ------------------------------------------------------------------------------------------------------------
def layer_norm(
    x: Tensor,
    weight: Optional[Tensor],
    bias: Optional[Tensor],
    eps: float = 1e-5,
    residual: Optional[Tensor] = None,
    out: Optional[Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    groups: int = 1,
) -> Tensor:
    """y = LayerNorm(x) (* weight + bias) (+ residual) over the last ``C = x.shape[1] / groups`` elements."""
    _need_cuda(x, weight, bias, residual, out)
    M, W, ldx = _rows(x)
    if W % groups:
        raise ValueError("row width not divisible by groups")
    C = W // groups
    if out is None:
        out = torch.empty((M, W), dtype=out_dtype or x.dtype, device=x.device)
    _, Wo, ldy = _rows(out)
    if Wo != W or out.shape[0] != M:
        raise ValueError("output shape mismatch")
    ldr, rdt = 0, F32
    if residual is not None:
        Mr, Wr, ldr = _rows(residual)
        if (Mr, Wr) != (M, W):
            raise ValueError("residual shape mismatch")
        rdt = dtype_code(residual.dtype)
    for p in (weight, bias):
        if p is not None and p.numel() != C:
            raise ValueError("LayerNorm parameter size mismatch")
    with _Timed("layer_norm", 8.0 * M * W, _nbytes(x, residual, out)):
        rc = _lib.load().anemoi_b200_layer_norm(
            _ptr(x), ldx, dtype_code(x.dtype), _ptr(_f32(weight)), _ptr(_f32(bias)), _ptr(residual), ldr, rdt, _ptr(out), ldy,
            dtype_code(out.dtype), M, groups, C, float(eps), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_layer_norm")
    return out


# ---- fp32 GEMMs on the tensor cores (exact bf16 x 3 split, csrc/misc.cu:split_bf16x3_kernel) -----------------------------------
# ANEMOI_B200_FP32_TC=0 keeps the FFMA kernel for every fp32 GEMM.
FP32_TC = os.environ.get("ANEMOI_B200_FP32_TC", "1") != "0"


def split_bf16x3(x: Tensor, weight_side: bool) -> Tensor:
    """fp32 [M, K] (K % 4 == 0) -> bf16 [M, 6K]: the operand (or weight) side of the six exact partial products."""
    _need_cuda(x)
    M, K, ldi = _rows(x)
    out = torch.empty((M, 6 * K), dtype=torch.bfloat16, device=x.device)
    with _Timed("split_bf16x3", 0.0, float(M) * K * 16):
        rc = _lib.load().anemoi_b200_split_bf16x3(_ptr(x), ldi, _ptr(out), 6 * K, M, K, 1 if weight_side else 0, _stream())
    _lib.check(rc, "anemoi_b200_split_bf16x3")
    return out


def _fp32_on_tensor_cores(a: Tensor, weight: Tensor) -> Optional[tuple]:
    """(a6, w6) when this fp32 GEMM should run as one bf16 tcgen05 GEMM over the split operands, else None (FFMA kernel)."""
    if not FP32_TC or a.dtype != torch.float32 or weight.dtype != torch.float32:
        return None
    M, K = a.shape
    N = weight.shape[0]
    if K % 4 or K < 16 or M < 256 or M * N * K < (1 << 24) or a.stride(0) % 4 or weight.stride(0) % 4 or a.data_ptr() % 16 or weight.data_ptr() % 16:
        return None
    # the weight is split on every call: it is small next to the activations, and a cache keyed on the storage address would go stale when
    # the allocator hands the same address to another tensor
    w6 = split_bf16x3(weight, True)
    return split_bf16x3(a, False), w6


def _linear_one(
    a: Tensor,
    weight: Tensor,
    bias: Optional[Tensor] = None,
    gelu: bool = False,
    residual: Optional[Tensor] = None,
    gather1: Optional[tuple[Tensor, Tensor]] = None,
    gather2: Optional[tuple[Tensor, Tensor]] = None,
    out: Optional[Tensor] = None,
    out_dtype: Optional[torch.dtype] = None,
    ln_stats: Optional[Tensor] = None,
    ln_colsum: Optional[Tensor] = None,
    ln_dim: int = 0,
    ln_eps: float = 0.0,
    stats_out: Optional[Tensor] = None,
    _nopdl: bool = False,
) -> Tensor:
    """out = [gelu](a @ weight.T + bias + g1[idx1] + g2[idx2]) + residual.

    ``stats_out`` (fp32 [M, ceil(N/64), 2], see ``partial_stats_buffer``) receives per-block (sum, sum of squares) of the stored output
    rows; a later ``linear`` consumes it as ``ln_stats`` of the same 3-D shape together with ``ln_dim`` (normalised width) and ``ln_eps``.

    With ``ln_stats`` [M, 2] (``row_stats(a)``) and ``ln_colsum`` [N] the LayerNorm of ``a`` is folded in:
    out = [gelu](rstd * (a @ weight.T - mean * colsum) + bias) for ``weight`` already scaled by the LayerNorm gamma.

    ``a`` [M, K] and ``weight`` [N, K] share a dtype: bf16 -> tcgen05 tensor cores; fp32 -> the same tensor cores on the exact bf16 x 3 split
    of both operands (``split_bf16x3``: fp32-grade result, ~1e-6), or the FFMA kernel for small / unaligned problems;
    ``gatherX = (table fp32 or bf16 [*, >=N], idx int32 [M])`` (bf16 tables halve the L2 traffic of the gather-add epilogue).
    """
    _need_cuda(a, weight, bias, residual, out)
    M, K, lda = _rows(a)
    N, Kw, ldw = _rows(weight)
    if K != Kw:
        raise ValueError(f"linear: inner dimensions differ ({K} vs {Kw})")
    if a.dtype != weight.dtype:
        raise TypeError(f"linear: operand dtypes differ ({a.dtype} vs {weight.dtype})")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype or a.dtype, device=a.device)
    Mo, No, ldo = _rows(out)
    if (Mo, No) != (M, N):
        raise ValueError("linear: output shape mismatch")
    ldr, rdt = 0, F32
    if residual is not None:
        Mr, Nr, ldr = _rows(residual)
        if (Mr, Nr) != (M, N):
            raise ValueError("linear: residual shape mismatch")
        rdt = dtype_code(residual.dtype)
    g1 = i1 = g2 = i2 = None
    ldg = 0
    for n, g in enumerate((gather1, gather2)):
        if g is None:
            continue
        tab, idx = g
        _need_cuda(tab, idx)
        _, Ng, ld = _rows(tab)
        if tab.dtype not in (torch.float32, torch.bfloat16) or Ng < N or idx.dtype != torch.int32 or idx.numel() != M or not idx.is_contiguous():
            raise TypeError("linear: gather tables must be float32 or bfloat16 [*, >=N] with contiguous int32 indices [M]")
        if ldg not in (0, ld):
            raise ValueError("linear: the two gather tables must share a leading dimension")
        ldg = ld
        if n == 0:
            g1, i1 = tab, idx
        else:
            g2, i2 = tab, idx
    if g1 is None and g2 is not None:
        g1, i1, g2, i2 = g2, i2, None, None
    gflags = (8 if g1 is not None and g1.dtype == torch.bfloat16 else 0) | (16 if g2 is not None and g2.dtype == torch.bfloat16 else 0)  # ANEMOI_EPI_G*_BF16
    if ln_stats is None and stats_out is None:
        split = _fp32_on_tensor_cores(a, weight)
        if split is not None:  # fp32 operands: ONE bf16 tcgen05 GEMM over the six exact partial products (inner dimension 6K)
            a, weight = split
            K, lda, ldw = 6 * K, 6 * K, 6 * K
    tc = a.dtype == torch.bfloat16 and K >= 64 and lda % 8 == 0 and ldw % 8 == 0 and a.data_ptr() % 16 == 0 and weight.data_ptr() % 16 == 0
    gbytes = (g1.element_size() * float(M) * N if g1 is not None else 0.0) + (g2.element_size() * float(M) * N if g2 is not None else 0.0)
    with _Timed("linear_tcgen05" if tc else "linear_ffma", 2.0 * M * N * K, _nbytes(a, weight, residual, out) + gbytes):
        if (ln_stats is None) != (ln_colsum is None):
            raise ValueError("linear: ln_stats and ln_colsum go together")
        ln_parts = 0
        if ln_stats is not None:
            if ln_stats.dtype != torch.float32 or not ln_stats.is_contiguous() or ln_colsum.numel() != N:
                raise ValueError("linear: ln_stats must be contiguous float32 and ln_colsum float32 [N]")
            if ln_stats.dim() == 3:  # partial (sum, sum of squares) per 64-column block from the producing GEMM
                ln_parts = ln_stats.shape[1]
                if ln_stats.shape != (M, ln_parts, 2) or ln_dim <= 0 or ln_parts != (ln_dim + 63) // 64:
                    raise ValueError("linear: partial ln_stats must be [M, ceil(ln_dim/64), 2] with ln_dim > 0")
            elif ln_stats.shape != (M, 2):
                raise ValueError("linear: ln_stats must be float32 [M, 2]")
        if stats_out is not None and (stats_out.dtype != torch.float32 or not stats_out.is_contiguous() or stats_out.shape != (M, (N + 63) // 64, 2)):
            raise ValueError("linear: stats_out must be contiguous float32 [M, ceil(N/64), 2]")
        _need_cuda(ln_stats, ln_colsum, stats_out)
        rc = _lib.load().anemoi_b200_linear(
            _ptr(a), lda, _ptr(weight), ldw, dtype_code(a.dtype), _ptr(_f32(bias)), _ptr(g1), _ptr(i1), _ptr(g2), _ptr(i2), ldg, _ptr(residual),
            ldr, rdt, _ptr(out), ldo, dtype_code(out.dtype), M, N, K, (EPI_GELU if gelu else 0) | _rev("gemm") | (4 if _nopdl else 0) | gflags, _ptr(ln_stats), _ptr(_f32(ln_colsum)),
            ln_parts, int(ln_dim), float(ln_eps), _ptr(stats_out), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_linear")
    return out


# ---- tail-wave split of narrow GEMMs ---------------------------------------------------------------------------------------------
# The CTA-pair GEMM walks 256 x 256 tiles on 74 pairs.  For N <= 512 a cfg2-sized problem (40 962 rows: 161 row tiles x 2) is 4.35 waves: the
# fifth wave keeps 26 of 74 pairs busy for a full tile time.  Split by rows instead: whole waves on the current stream, the remaining rows
# as a second launch on a forked stream (it only depends on what the first depends on; its small problem takes the 128 x 128 single-CTA
# tiles, half the per-SM work), whose CTAs fill the SMs as the first launch's CTAs retire; the current stream joins before going on.
# Graph-capturable (fork / join through events).  Measured (profiles/r2/call45_ab_tail_split.txt, same call): SLOWER - projection 43.0 -> 54.6 us,
# MLP-2 80.7 -> 84.5 us, cfg2 step 8.43 -> 8.54 ms: the second launch's fill and the fork / join cost more than the idle pairs.  Opt-in
# (ANEMOI_B200_TAIL_SPLIT=1), off by default.
TAIL_SPLIT = os.environ.get("ANEMOI_B200_TAIL_SPLIT", "0") != "0"
_SIDE = {}


def _tail_split_rows(a: Tensor, weight: Tensor, kw: dict) -> int:
    """Rows of the whole-wave part when this GEMM should be split (0 = run it as one launch)."""
    if not TAIL_SPLIT or a.dtype != torch.bfloat16 or weight.shape[0] > 512 or kw.get("stats_out") is not None:
        return 0
    M, N = a.shape[0], weight.shape[0]
    pairs = torch.cuda.get_device_properties(a.device).multi_processor_count // 2
    tiles_n = -(-N // 256)
    row_tiles = -(-M // 256)
    if row_tiles * tiles_n < 2 * pairs:  # not the CTA-pair kernel (linear_tcgen05's own rule)
        return 0
    per_wave = pairs // tiles_n  # row tiles per full wave
    full = row_tiles // per_wave
    rest = row_tiles - full * per_wave
    if full < 2 or rest == 0 or rest * tiles_n > 0.6 * pairs:  # a tail wave that is more than ~60 % full is not worth a second launch
        return 0
    return full * per_wave * 256


def linear(a: Tensor, weight: Tensor, bias: Optional[Tensor] = None, **kw) -> Tensor:
    """``_linear_one`` (see there for the arguments), with the tail wave of a narrow GEMM launched on a forked stream."""
    rows = _tail_split_rows(a, weight, kw) if a.is_cuda else 0
    if not rows:
        return _linear_one(a, weight, bias, **kw)
    M, N = a.shape[0], weight.shape[0]
    out = kw.pop("out", None)
    if out is None:
        out = torch.empty((M, N), dtype=kw.get("out_dtype") or a.dtype, device=a.device)

    def part(sl):
        k2 = dict(kw)
        for name in ("residual", "ln_stats"):
            if k2.get(name) is not None:
                k2[name] = k2[name][sl]
        for name in ("gather1", "gather2"):
            if k2.get(name) is not None:
                k2[name] = (k2[name][0], k2[name][1][sl])
        return _linear_one(a[sl], weight, bias, out=out[sl], _nopdl=sl.start != 0, **k2)  # the forked launch follows an event wait: plain launch

    cur = torch.cuda.current_stream(a.device)
    side = _SIDE.get(a.device.index)
    if side is None:
        side = _SIDE[a.device.index] = torch.cuda.Stream(a.device)
    fork = torch.cuda.Event()
    fork.record(cur)
    part(slice(0, rows))
    side.wait_event(fork)
    with torch.cuda.stream(side):
        part(slice(rows, M))
        join = torch.cuda.Event()
        join.record(side)
    cur.wait_event(join)
    return out


ATTN_MAX_EDGE_DIM = 16  # attributes per edge the fused attention kernel accepts (rows zero-padded to this many floats)


def attention_tiles_supported(channels: int, heads: int, dtype: torch.dtype, dp: int) -> bool:
    """Shapes the destination-tile kernel handles: bf16, 32 or 64 channels per head, heads a multiple of 256 / Ch, dp in {4, 8, 12, 16}."""
    if dtype != torch.bfloat16 or heads <= 0 or channels % heads:
        return False
    ch = channels // heads
    return ch in (32, 64) and heads % (256 // ch) == 0 and dp in (4, 8, 12, 16)


def attention_fold_supported(channels: int, heads: int, dtype: torch.dtype, edge_dim: int) -> bool:
    """True if the coalesced slab attention kernel (and with it the folded lin_edge form) handles this shape."""
    epc = 16 // (2 if dtype == torch.bfloat16 else 4)
    if channels % heads or channels % epc or edge_dim > ATTN_MAX_EDGE_DIM:
        return False
    ch = channels // heads
    if ch % epc or ch // epc not in (2, 4, 8, 16):
        return False
    chunks = channels // epc
    return chunks < 32 or chunks % 32 == 0


def gt_attention(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    csr: GraphCSR,
    heads: int,
    e_proj: Optional[Tensor] = None,
    edge_attr: Optional[Tensor] = None,
    w_edge: Optional[Tensor] = None,
    b_edge: Optional[Tensor] = None,
    qw: Optional[Tensor] = None,
    abar: Optional[Tensor] = None,
    dp: int = 0,
    add: Optional[Tensor] = None,
    out: Optional[Tensor] = None,
    tiles: Optional[AttnTilePlan] = None,
    lse: Optional[Tensor] = None,
) -> Tensor:
    """Edge-softmax attention over the cached CSR.  q [n_dst, H*Ch]; k, v [n_src, H*Ch] (column slices allowed).
    ``tiles`` (folded bf16 form only): run the destination-tile tensor-core kernel on this plan of the same CSR.

    Edge term, one of: ``e_proj`` [E, H*Ch] (materialised lin_edge output, the reference operator boundary);
    ``edge_attr`` fp32 [E, >=d_e] + ``w_edge`` fp32 [H*Ch, d_e] (+ ``b_edge``): projection inside the kernel;
    ``edge_attr`` fp32 [E, 16] (zero-padded) + ``qw`` / ``abar`` [n_dst, >= H*dp]: folded form (see include/anemoi_b200.h).
    ``add`` is summed into the output.
    """
    _need_cuda(q, k, v, e_proj, edge_attr, w_edge, b_edge, qw, abar, add, out)
    n_dst, C, ldq = _rows(q)
    n_src, Ck, ldk = _rows(k)
    n_src_v, Cv, ldv = _rows(v)
    if not (C == Ck == Cv) or n_src != n_src_v or C % heads:
        raise ValueError("gt_attention: q/k/v shape mismatch")
    if n_dst != csr.n_dst or n_src != csr.n_src:
        raise ValueError(f"gt_attention: CSR is for ({csr.n_src}, {csr.n_dst}) nodes, tensors have ({n_src}, {n_dst})")
    if not (q.dtype == k.dtype == v.dtype):
        raise TypeError("gt_attention: q/k/v dtypes differ")
    if out is None:
        out = torch.empty((n_dst, C), dtype=q.dtype, device=q.device)
    _, Co, ldo = _rows(out)
    if Co != C or out.dtype != q.dtype:
        raise ValueError("gt_attention: output mismatch")
    lde_proj = lde = ldw_e = d_e = ldadd = ldqw = ldabar = 0
    if e_proj is not None:
        E, Ce, lde_proj = _rows(e_proj)
        if E != csr.n_edges or Ce != C or e_proj.dtype != q.dtype:
            raise ValueError("gt_attention: e_proj mismatch")
    if edge_attr is not None:
        E, d_e_pad, lde = _rows(edge_attr)
        if E != csr.n_edges or edge_attr.dtype != torch.float32:
            raise ValueError("gt_attention: edge_attr must be float32 [E, >= d_e]")
        if qw is not None:
            if abar is None or dp <= 0:
                raise ValueError("gt_attention: folded form needs qw, abar and dp")
            _, wq, ldqw = _rows(qw)
            _, wa, ldabar = _rows(abar)
            if wq < heads * dp or wa < heads * dp or qw.dtype != q.dtype or abar.dtype != q.dtype or qw.shape[0] != n_dst or abar.shape[0] != n_dst:
                raise ValueError("gt_attention: qw / abar mismatch")
            d_e = min(dp, d_e_pad)
        else:
            Cw, d_e, ldw_e = _rows(w_edge)
            if Cw != C or d_e > d_e_pad or w_edge.dtype != torch.float32:
                raise ValueError("gt_attention: fused lin_edge arguments mismatch")
    if add is not None:
        na, Ca, ldadd = _rows(add)
        if (na, Ca) != (n_dst, C) or add.dtype != q.dtype:
            raise ValueError("gt_attention: add mismatch")
    # algorithmic bytes: q, k, v, add read once + out written once + per-edge index/attributes (or e_proj)
    es = q.element_size()
    abytes = es * C * (2.0 * n_dst + 2.0 * n_src + (n_dst if add is not None else 0)) + csr.n_edges * (4.0 + 4.0 * lde + (es * C if e_proj is not None else 0)) + 4.0 * n_dst
    aflops = csr.n_edges * (4.0 * C + 4.0 * d_e * heads) + n_dst * 4.0 * d_e * C
    if csr.n_edges == 0 and abar is not None:
        # an empty edge tensor has a null data pointer, so the library sees "no edge term" and leaves abar untouched; with no edges
        # abar = sum_e alpha_e a_e is exactly zero (tests/test_gpu_parity.py::test_edge_cases_empty_and_isolated)
        abar.zero_()
    if lse is not None:
        _need_cuda(lse)
        if qw is not None or tiles is not None or lse.dtype != torch.float32 or not lse.is_contiguous() or tuple(lse.shape) != (n_dst, heads):
            raise ValueError("gt_attention: lse must be contiguous float32 [n_dst, heads] and goes with the materialised / in-kernel projection forms")
    if tiles is not None:
        if qw is None or e_proj is not None or q.dtype != torch.bfloat16:
            raise ValueError("gt_attention: the tiled kernel implements the folded bf16 form (qw / abar)")
        if lde < dp or tiles.max_edges > min(224, 10752 // (4 * dp)):
            raise ValueError("gt_attention: edge_attr rows narrower than dp, or a tile plan built for a smaller dp")
        with _Timed("gt_attention", aflops, abytes):
            rc = _lib.load().anemoi_b200_gt_attention_tiled_fwd(
                _ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(edge_attr), lde, _ptr(_f32(b_edge)), _ptr(qw), ldqw, _ptr(abar), ldabar, dp,
                _ptr(csr.colptr32), _ptr(tiles.tile_meta), _ptr(tiles.slot_src), _ptr(tiles.emeta), tiles.n_tiles, _ptr(add), ldadd, _ptr(out), ldo,
                n_dst, heads, C // heads, _stream())  # fmt: skip
        _lib.check(rc, "anemoi_b200_gt_attention_tiled_fwd")
        return out
    with _Timed("gt_attention", aflops, abytes):
        rc = _lib.load().anemoi_b200_gt_attention_fwd(
            _ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(e_proj), lde_proj, _ptr(edge_attr), lde, d_e, _ptr(w_edge), ldw_e, _ptr(_f32(b_edge)),
            _ptr(qw), ldqw, _ptr(abar), ldabar, dp, _ptr(csr.src32) or _ptr(csr.colptr32), _ptr(csr.colptr32), _ptr(add), ldadd, _ptr(out), ldo, _ptr(lse),
            n_dst, heads, C // heads, dtype_code(q.dtype), _rev("gemm"), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_gt_attention_fwd")
    return out


def graphconv_ln_aggregate(
    h: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], e: Tensor, csr: GraphCSR, eps: float = 1e-5, out: Optional[Tensor] = None
) -> tuple[Tensor, Tensor]:
    """(e_new, out): e_new = LayerNorm(h) + e ; out[d] = sum of e_new over the (dst-sorted) edges into d."""
    _need_cuda(h, weight, bias, e, out)
    E, C, ldh = _rows(h)
    Ee, Ce, lde = _rows(e)
    if (E, C) != (Ee, Ce) or E != csr.n_edges or h.dtype != e.dtype:
        raise ValueError("graphconv_ln_aggregate: h / e mismatch")
    e_new = torch.empty((E, C), dtype=h.dtype, device=h.device)
    if out is None:
        out = torch.empty((csr.n_dst, C), dtype=h.dtype, device=h.device)
    no, Co, ldo = _rows(out)
    if (no, Co) != (csr.n_dst, C) or out.dtype != h.dtype:
        raise ValueError("graphconv_ln_aggregate: output mismatch")
    with _Timed("graphconv_ln_aggregate", 10.0 * E * C, _nbytes(h, e, e_new, out) + 4.0 * csr.n_dst):
        rc = _lib.load().anemoi_b200_graphconv_ln_aggregate(
            _ptr(h), ldh, _ptr(_f32(weight)), _ptr(_f32(bias)), _ptr(e), lde, _ptr(e_new), C, _ptr(csr.colptr32), _ptr(out), ldo, csr.n_dst, C,
            float(eps), dtype_code(h.dtype), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_graphconv_ln_aggregate")
    return e_new, out


GRAPHCONV_FUSED_WIDTHS = (16, 32, 64)  # widths the one-kernel GraphConv is built for (csrc/graphconv_fused.cu)
GRAPHCONV_FUSED_MAX_LAYERS = 6


def graphconv_fused(
    x_src: Tensor, x_dst: Tensor, e: Tensor, weights: Tensor, biases: Tensor, n_layers: int, gamma: Optional[Tensor], beta: Optional[Tensor],
    csr: GraphCSR, eps: float = 1e-5, out: Optional[Tensor] = None,
) -> tuple[Tensor, Tensor]:  # fmt: skip
    """The whole GraphConv operator (layers/conv.py:66-81) as ONE kernel, C in {16, 32, 64}:
    (e_new, out) with e_new = LayerNorm(edge_mlp([x_dst[dst]; x_src[src]; e])) + e and out[d] = sum of e_new over the edges into d.
    ``weights``: packed [C, 3C] + (n_layers - 1) x [C, C] of the compute dtype; ``biases``: fp32 [n_layers, C]."""
    _need_cuda(x_src, x_dst, e, weights, biases, gamma, beta, out)
    E, C, lde = _rows(e)
    ns, Cs, lds = _rows(x_src)
    nd, Cd, ldd = _rows(x_dst)
    dt = e.dtype
    if Cs != C or Cd != C or E != csr.n_edges or nd != csr.n_dst or x_src.dtype != dt or x_dst.dtype != dt or weights.dtype != dt:
        raise ValueError("graphconv_fused: operand mismatch")
    if weights.numel() != (n_layers + 2) * C * C or not weights.is_contiguous() or biases.numel() != n_layers * C or biases.dtype != torch.float32:
        raise ValueError("graphconv_fused: packed weights / biases do not match n_layers and C")
    e_new = torch.empty((E, C), dtype=dt, device=e.device)
    if out is None:
        out = torch.empty((csr.n_dst, C), dtype=dt, device=e.device)
    no, Co, ldo = _rows(out)
    if (no, Co) != (csr.n_dst, C) or out.dtype != dt:
        raise ValueError("graphconv_fused: output mismatch")
    with _Timed("graphconv_fused", 2.0 * (n_layers + 2) * C * C * E, _nbytes(e, e_new, out) + 12.0 * E):
        rc = _lib.load().anemoi_b200_graphconv_fused(
            _ptr(x_src), lds, _ptr(x_dst), ldd, _ptr(e), lde, _ptr(weights), _ptr(biases.contiguous()), n_layers, _ptr(_f32(gamma)), _ptr(_f32(beta)),
            _ptr(e_new), C, _ptr(csr.src32), _ptr(csr.dst32), _ptr(csr.colptr32), _ptr(out), ldo, csr.n_dst, E, C, float(eps), dtype_code(dt),
            _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_graphconv_fused")
    return e_new, out


def cast_pad(x: Tensor, dtype: torch.dtype, k_pad: Optional[int] = None, idx: Optional[Tensor] = None, out: Optional[Tensor] = None) -> Tensor:
    """Copy/cast [M, K] into [M, k_pad] (zero-filled tail), optionally gathering rows ``x[idx]`` (idx int32)."""
    _need_cuda(x, idx, out)
    M, K, ldi = _rows(x)
    if idx is not None:
        if idx.dtype != torch.int32 or not idx.is_contiguous():
            raise TypeError("cast_pad: idx must be contiguous int32")
        M = idx.numel()
    k_pad = K if k_pad is None else k_pad
    if out is None:
        out = torch.empty((M, k_pad), dtype=dtype, device=x.device)
    Mo, Ko, ldo = _rows(out)
    if Mo != M or Ko < k_pad:
        raise ValueError("cast_pad: output mismatch")
    with _Timed("cast_pad", 0.0, float(M) * (K * x.element_size() + k_pad * out.element_size())):
        rc = _lib.load().anemoi_b200_cast_pad(_ptr(x), ldi, dtype_code(x.dtype), _ptr(idx), _ptr(out), ldo, dtype_code(out.dtype), M, K, k_pad, _stream())
    _lib.check(rc, "anemoi_b200_cast_pad")
    return out


def add(a: Tensor, b: Tensor, out_dtype: Optional[torch.dtype] = None) -> Tensor:
    """a + b (fp32 add, any mix of f32/bf16 operands) — the latent skip around the processor."""
    _need_cuda(a, b)
    M, C, lda = _rows(a)
    Mb, Cb, ldb = _rows(b)
    if (M, C) != (Mb, Cb):
        raise ValueError("add: shape mismatch")
    out = torch.empty((M, C), dtype=out_dtype or a.dtype, device=a.device)
    with _Timed("add", float(M) * C, _nbytes(a, b, out)):
        rc = _lib.load().anemoi_b200_add(_ptr(a), lda, dtype_code(a.dtype), _ptr(b), ldb, dtype_code(b.dtype), _ptr(out), C, dtype_code(out.dtype), M, C, _stream())
    _lib.check(rc, "anemoi_b200_add")
    return out


def partial_stats_buffer(out_rows: int, out_cols: int, device) -> Tensor:
    """Buffer for ``linear(..., stats_out=)``: [M, ceil(N/64), 2] fp32."""
    return torch.empty((out_rows, (out_cols + 63) // 64, 2), dtype=torch.float32, device=device)


def row_stats(x: Tensor, eps: float = 1e-5) -> Tensor:
    """Per-row LayerNorm statistics [M, 2] = (mean, 1/sqrt(var + eps)) for the LayerNorm folded into ``linear``."""
    _need_cuda(x)
    M, C, ldx = _rows(x)
    stats = torch.empty((M, 2), dtype=torch.float32, device=x.device)
    with _Timed("row_stats", 4.0 * M * C, float(M) * C * x.element_size() + 8.0 * M):
        rc = _lib.load().anemoi_b200_row_stats(_ptr(x), ldx, dtype_code(x.dtype), _ptr(stats), M, C, float(eps), _rev("stats"), _stream())
    _lib.check(rc, "anemoi_b200_row_stats")
    return stats


def assemble_input(x: Tensor, attrs: Optional[Tensor], out_dtype: torch.dtype, k_pad: Optional[int] = None) -> Tensor:
    """``cat([rearrange(x, "b t e g v -> (b e g) (t v)"), attrs], -1)`` in ``out_dtype``; with ``k_pad`` the rows are zero-padded to that
    width, so the result is directly the (K-padded) A operand of the embedding GEMM.  Returns [B*E*G, k_pad or T*V+A].
    Reference: models/encoder_processor_decoder.py:98-127."""
    _need_cuda(x, attrs)
    if x.dim() != 5:
        raise ValueError("assemble_input: x must be (batch, time, ensemble, grid, vars)")
    x = x.contiguous().float()
    B, T, E, G, V = x.shape
    A = 0 if attrs is None else attrs.shape[1]
    if attrs is not None:
        attrs = attrs.contiguous().float()
        if B * E * G > 0 and (attrs.shape[0] == 0 or (B * E * G) % attrs.shape[0] != 0):
            raise ValueError(f"assemble_input: {attrs.shape[0]} attribute rows do not tile {B * E * G} node rows")
    K = T * V + A
    Kpad = K if k_pad is None else int(k_pad)
    if Kpad < K:
        raise ValueError("assemble_input: k_pad smaller than the assembled width")
    out = torch.empty((B * E * G, Kpad), dtype=out_dtype, device=x.device)
    with _Timed("assemble_input", 0.0, float(x.numel()) * 4 + out.numel() * out.element_size()):
        rc = _lib.load().anemoi_b200_assemble_input(_ptr(x), B, T, E, G, V, _ptr(attrs), A, attrs.shape[0] if attrs is not None else 0, _ptr(out),
                                                    Kpad, Kpad, dtype_code(out_dtype), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_assemble_input")
    return out


def assemble_output(dec: Tensor, x: Optional[Tensor], batch: int, ensemble: int, n_step_output: int, step: int = -1,
                    skip_src: Optional[Tensor] = None, bound: Optional[Tensor] = None) -> Tensor:  # fmt: skip
    """``rearrange(dec, "(b e g) (t v) -> b t e g v")`` as fp32 + residual of the input's ``step`` slice on the prognostic variables +
    per-variable ReLU / LeakyReLU bounding, one pass.  ``skip_src`` / ``bound``: int32 [V_out] device tensors (see include/anemoi_b200.h).
    Reference: models/encoder_processor_decoder.py:129-163."""
    _need_cuda(dec)
    M, C, ldd = _rows(dec)
    if M % (batch * ensemble) or C % n_step_output:
        raise ValueError("assemble_output: decoder output does not factor into (batch ensemble grid) x (time vars)")
    G, V_out = M // (batch * ensemble), C // n_step_output
    T_in = V_in = 0
    if x is not None:
        _need_cuda(x)
        x = x.contiguous().float()
        T_in, V_in = x.shape[1], x.shape[4]
        if x.shape[0] != batch or x.shape[2] != ensemble or x.shape[3] != G:
            raise ValueError("assemble_output: residual input does not match the decoder output rows")
        step = step % T_in
    y = torch.empty((batch, n_step_output, ensemble, G, V_out), dtype=torch.float32, device=dec.device)
    with _Timed("assemble_output", 0.0, float(M) * C * dec.element_size() + y.numel() * 4.0):
        rc = _lib.load().anemoi_b200_assemble_output(_ptr(dec), ldd, dtype_code(dec.dtype), _ptr(x) if x is not None else None, batch, T_in, ensemble, G,
                                                     V_in, step, _ptr(skip_src) if skip_src is not None else None,
                                                     _ptr(bound) if bound is not None else None, _ptr(y), n_step_output, V_out, _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_assemble_output")
    return y


GLU_ACTS = {"glu": 0, "swiglu": 1, "geglu": 2, "reglu": 3}


def glu_combine(gv: Tensor, act: str) -> Tensor:
    """``act(gv[:, :H]) * gv[:, H:]`` for the gated MLP variants (reference layers/mlp.py:38-53); ``gv`` = [gate_proj(x) | value_proj(x)]."""
    _need_cuda(gv)
    M, W, ldi = _rows(gv)
    if W % 2 or act not in GLU_ACTS:
        raise ValueError(f"glu_combine: need an even width and act in {sorted(GLU_ACTS)}")
    H = W // 2
    out = torch.empty((M, H), dtype=gv.dtype, device=gv.device)
    with _Timed("glu_combine", 8.0 * M * H, float(M) * 3 * H * gv.element_size()):
        rc = _lib.load().anemoi_b200_glu_combine(_ptr(gv), ldi, _ptr(out), H, M, H, GLU_ACTS[act], dtype_code(gv.dtype), _stream())
    _lib.check(rc, "anemoi_b200_glu_combine")
    return out


def glu_combine_bwd(gv: Tensor, dy: Tensor, act: str) -> Tensor:
    """Gradient of ``glu_combine`` with respect to ``gv`` = [gate | value] for the cotangent ``dy`` of its output."""
    _need_cuda(gv, dy)
    M, W, ldi = _rows(gv)
    H = W // 2
    Md, Hd, lddy = _rows(dy)
    if W % 2 or act not in GLU_ACTS or (Md, Hd) != (M, H) or dy.dtype != gv.dtype:
        raise ValueError("glu_combine_bwd: operand mismatch")
    dgv = torch.empty((M, W), dtype=gv.dtype, device=gv.device)
    with _Timed("glu_combine_bwd", 16.0 * M * H, float(M) * 5 * H * gv.element_size()):
        rc = _lib.load().anemoi_b200_glu_combine_bwd(_ptr(gv), ldi, _ptr(dy), lddy, _ptr(dgv), W, M, H, GLU_ACTS[act], dtype_code(gv.dtype), _stream())
    _lib.check(rc, "anemoi_b200_glu_combine_bwd")
    return dgv


def cond_layer_norm(x: Tensor, cond: Tensor, w_scale: Tensor, b_scale: Tensor, w_bias: Tensor, b_bias: Tensor, eps: float = 1e-5,
                    out_dtype: Optional[torch.dtype] = None) -> Tensor:  # fmt: skip
    """``LN(x) * (1 + cond @ w_scale.T + b_scale) + (cond @ w_bias.T + b_bias)`` (reference ConditionalLayerNorm, normalization.py:34-94)."""
    _need_cuda(x, cond, w_scale, b_scale, w_bias, b_bias)
    M, C, ldx = _rows(x)
    cond = cond.float().contiguous() if cond.dtype != torch.float32 or not cond.is_contiguous() else cond
    Mc, Dc, ldc = _rows(cond)
    if Mc != M or tuple(w_scale.shape) != (C, Dc) or tuple(w_bias.shape) != (C, Dc) or b_scale.numel() != C or b_bias.numel() != C:
        raise ValueError("cond_layer_norm: shape mismatch (cond [M, Dc], weights [C, Dc], biases [C])")
    out = torch.empty((M, C), dtype=out_dtype or x.dtype, device=x.device)
    with _Timed("cond_layer_norm", (8.0 + 4.0 * Dc) * M * C, _nbytes(x, out)):
        rc = _lib.load().anemoi_b200_cond_layer_norm(_ptr(x), ldx, dtype_code(x.dtype), _ptr(cond), ldc, _ptr(_f32(w_scale)), _ptr(_f32(b_scale)),
                                                     _ptr(_f32(w_bias)), _ptr(_f32(b_bias)), _ptr(out), C, dtype_code(out.dtype), M, C, Dc, float(eps),
                                                     _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_cond_layer_norm")
    return out


# ------------------------------------------------------------------------------------------------------------
# multi-GPU exchange over peer memory (csrc/peer.cu); ``ch`` is a distributed.peer.PeerChannel
# ------------------------------------------------------------------------------------------------------------
def peer_rendezvous(ch, like: Tensor) -> None:
    """Start exchange c on this rank: announce it to the peers and wait until all of them have started it too."""
    _need_cuda(like)
    with _Timed("peer_rendezvous"):
        rc = _lib.load().anemoi_b200_peer_rendezvous(ch.ctl_ptrs, ch.world, ch.rank, _stream())
    _lib.check(rc, "anemoi_b200_peer_rendezvous")


def halo_push(ch, rows: Tensor, send_idx: Tensor, send_off: Tensor, n_send: int, dst_ptrs, dst_ld_bytes: int) -> None:
    """rows[send_idx[j]] -> the peers' tables (``dst_ptrs``: ctypes uint64 [world] of peer-mapped addresses), then signal arrival."""
    _need_cuda(rows, send_idx, send_off)
    _, W, ld = _rows(rows)
    es = rows.element_size()
    if send_idx.dtype != torch.int32 or send_off.dtype != torch.int32 or send_off.numel() != ch.world + 1:
        raise TypeError("halo_push: send_idx / send_off must be int32 (send_off has world + 1 entries)")
    with _Timed("halo_push", 0.0, 2.0 * n_send * W * es):
        rc = _lib.load().anemoi_b200_halo_push(_ptr(rows), ld * es, W * es, _ptr(send_idx), _ptr(send_off), n_send, dst_ptrs, dst_ld_bytes, ch.ctl_ptrs,
                                               ch.world, ch.rank, _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_halo_push")


def halo_wait(ch, like: Tensor) -> None:
    """Wait (on the stream) until every peer's rows of the current exchange have arrived."""
    _need_cuda(like)
    with _Timed("halo_wait"):
        rc = _lib.load().anemoi_b200_halo_wait(ch.ctl_ptrs, ch.world, ch.rank, _stream())
    _lib.check(rc, "anemoi_b200_halo_wait")


# ------------------------------------------------------------------------------------------------------------
# backward of the fused ops (csrc/backward.cu)
# ------------------------------------------------------------------------------------------------------------
def reverse_csr(csr: GraphCSR) -> tuple[Tensor, Tensor]:
    """(rev_ptr32 [n_src + 1], rev_eid32 [E]): the edge ids stably sorted by source - the reverse structure of the reference's
    ``edge_index_to_csc(..., reverse=True)`` (triton/utils.py:61-68: ``argsort(row, stable=True)``), built once per graph with the same
    torch integer ops and cached on the CSR."""
    if csr.rev is None:
        src = csr.src32.long()
        eid = torch.sort(src, stable=True).indices.to(torch.int32).contiguous()
        ptr = torch.zeros(csr.n_src + 1, dtype=torch.int64, device=src.device)
        ptr[1:] = torch.cumsum(torch.bincount(src, minlength=csr.n_src), 0)
        csr.rev = (ptr.to(torch.int32).contiguous(), eid)
    return csr.rev


def gt_attention_bwd(q: Tensor, k: Tensor, v: Tensor, e_proj: Optional[Tensor], out: Tensor, dout: Tensor, lse: Tensor, csr: GraphCSR, heads: int):
    """(dq, dk, dv, de) of ``gt_attention(q, k, v, csr, heads, e_proj=e_proj)`` for the cotangent ``dout``; ``out`` / ``lse`` from the forward."""
    _need_cuda(q, k, v, e_proj, out, dout, lse)
    n_dst, C, ldq = _rows(q)
    n_src, _, ldk = _rows(k)
    _, _, ldv = _rows(v)
    _, _, ldo = _rows(out)
    _, _, lddo = _rows(dout)
    if not (q.dtype == k.dtype == v.dtype == out.dtype == dout.dtype) or lse.dtype != torch.float32 or not lse.is_contiguous():
        raise TypeError("gt_attention_bwd: q / k / v / out / dout share a dtype, lse is contiguous float32")
    rev_ptr, rev_eid = reverse_csr(csr)
    dq, dk, dv = torch.empty((n_dst, C), dtype=q.dtype, device=q.device), torch.empty((n_src, C), dtype=q.dtype, device=q.device), torch.empty((n_src, C), dtype=q.dtype, device=q.device)
    de, lde, ldde = None, 0, 0
    if e_proj is not None:
        _, _, lde = _rows(e_proj)
        de = torch.empty((csr.n_edges, C), dtype=q.dtype, device=q.device)
        ldde = C
    alpha = torch.empty((csr.n_edges, heads), dtype=torch.float32, device=q.device)
    ds = torch.empty_like(alpha)
    es = q.element_size()
    with _Timed("gt_attention_bwd", csr.n_edges * 10.0 * C, es * C * (4.0 * n_dst + 4.0 * n_src + 4.0 * csr.n_edges)):
        rc = _lib.load().anemoi_b200_gt_attention_bwd(
            _ptr(q), ldq, _ptr(k), ldk, _ptr(v), ldv, _ptr(e_proj), lde, _ptr(out), ldo, _ptr(dout), lddo, _ptr(lse), _ptr(csr.src32) or _ptr(csr.colptr32),
            _ptr(csr.colptr32), _ptr(csr.dst32) or _ptr(csr.colptr32), _ptr(rev_ptr), _ptr(rev_eid) or _ptr(rev_ptr), _ptr(dq), C, _ptr(dk), C, _ptr(dv), C,
            _ptr(de), ldde, _ptr(alpha) or _ptr(lse), _ptr(ds) or _ptr(lse), n_src, n_dst, heads, C // heads, dtype_code(q.dtype), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_gt_attention_bwd")
    return dq, dk, dv, de


def segment_sum(rows: Tensor, ptr32: Tensor, eid32: Optional[Tensor], n_out: int) -> Tensor:
    """``out[n] = sum_{j in [ptr32[n], ptr32[n+1])} rows[eid32[j] if eid32 is not None else j]`` (fp32 accumulation, deterministic): the backward
    of a row gather over a sorted index list / its reverse CSR."""
    _need_cuda(rows, ptr32, eid32)
    _, C, ld = _rows(rows)
    out = torch.empty((n_out, C), dtype=rows.dtype, device=rows.device)
    with _Timed("segment_sum", 1.0 * rows.shape[0] * C, _nbytes(rows, out)):
        rc = _lib.load().anemoi_b200_segment_sum(_ptr(rows), ld, _ptr(ptr32), _ptr(eid32), _ptr(out), C, n_out, C, dtype_code(rows.dtype), _stream())
    _lib.check(rc, "anemoi_b200_segment_sum")
    return out


_COL_SUM_CHUNKS = 256  # partial rows of the two-stage column sum


def col_sum(x: Tensor) -> Tensor:
    """``x.sum(0)`` in fp32 (deterministic two-stage reduction): the bias gradient of a Linear from its cotangent."""
    _need_cuda(x)
    M, N, ld = _rows(x)
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    partial = torch.empty((_COL_SUM_CHUNKS, (N + 3) // 4 * 4), dtype=torch.float32, device=x.device)
    with _Timed("col_sum", 1.0 * M * N, _nbytes(x, out)):
        rc = _lib.load().anemoi_b200_col_sum(_ptr(x), ld, _ptr(out), _ptr(partial), _COL_SUM_CHUNKS, M, N, dtype_code(x.dtype), _stream())
    _lib.check(rc, "anemoi_b200_col_sum")
    return out


_LN_BWD_BLOCKS = 592  # partial-sum rows of dgamma / dbeta: 4 CTAs per SM
LN_BWD_COL_SUM = os.environ.get("ANEMOI_B200_LN_BWD_COL_SUM", "1") != "0"


def layer_norm_bwd(x: Tensor, gamma: Optional[Tensor], dy: Optional[Tensor], eps: float, groups: int = 1, dz: Optional[Tensor] = None,
                   idx: Optional[Tensor] = None, want_dres: bool = False):  # fmt: skip
    """Backward of ``y = LayerNorm(x) * gamma + beta`` over ``groups`` groups of C channels per row for the cotangent
    ``g[r] = dy[r] + dz[idx[r]]``: returns (dx, dgamma [C], dbeta [C], g if ``want_dres``)."""
    _need_cuda(x, gamma, dy, dz, idx)
    M, W, ldx = _rows(x)
    C = W // groups
    dx = torch.empty((M, W), dtype=x.dtype, device=x.device)
    dres = torch.empty((M, W), dtype=x.dtype, device=x.device) if want_dres else None
    partial = torch.empty((_LN_BWD_BLOCKS, 2, C), dtype=torch.float32, device=x.device)
    lddy = lddz = 0
    if dy is not None:
        _, _, lddy = _rows(dy)
    if dz is not None:
        _, _, lddz = _rows(dz)
        if idx is not None and (idx.dtype != torch.int32 or idx.numel() != M * groups):
            raise TypeError("layer_norm_bwd: idx must be int32 with one entry per normalised row")
    with _Timed("layer_norm_bwd", 16.0 * M * W, _nbytes(x, dy, dx, dres)):
        rc = _lib.load().anemoi_b200_layer_norm_bwd(_ptr(x), ldx, _ptr(_f32(gamma)), _ptr(dy), lddy, _ptr(dz), lddz, _ptr(idx), _ptr(dx), W, _ptr(dres), W,
                                                    _ptr(partial), _LN_BWD_BLOCKS, M, groups, C, float(eps), dtype_code(x.dtype), _stream())  # fmt: skip
    _lib.check(rc, "anemoi_b200_layer_norm_bwd")
    # dgamma | dbeta = column sums of the per-CTA partial rows (two tiny launches of the column-sum kernel instead of a PyTorch reduction)
    sums = col_sum(partial.view(_LN_BWD_BLOCKS, 2 * C)).view(2, C) if (LN_BWD_COL_SUM and C % 2 == 0) else partial.sum(0)
    return dx, sums[0], sums[1], dres


GELU_BWD_FAST = os.environ.get("ANEMOI_B200_GELU_BWD_FAST", "1") != "0"  # bf16 cotangents: two-MUFU gelu' (|error| <= 1.3e-5) instead of erff + expf


def gelu(x: Tensor, dy: Optional[Tensor] = None) -> Tensor:
    """``gelu(x)`` (exact erf), or with ``dy`` the backward ``dy * gelu'(x)``."""
    _need_cuda(x, dy)
    M, N, ldx = _rows(x)
    y = torch.empty((M, N), dtype=x.dtype, device=x.device)
    lddy = 0
    if dy is not None:
        _, _, lddy = _rows(dy)
    mode = 0 if dy is None else 2 if (GELU_BWD_FAST and x.dtype == torch.bfloat16) else 1
    with _Timed("gelu", 10.0 * M * N, _nbytes(x, dy, y)):
        rc = _lib.load().anemoi_b200_gelu(_ptr(x), ldx, _ptr(dy), lddy, _ptr(y), N, M, N, mode, dtype_code(x.dtype), _stream())
    _lib.check(rc, "anemoi_b200_gelu")
    return y
