"""Cycle timeline of ONE launch of the tcgen05 GEMM (CTA 0), from a -DANEMOI_GEMM_TIMELINE build of the library:
    make -C anemoi_core_b200/csrc BUILD=build_tl LIB=../lib/variants/gemm_timeline.so EXTRA=-DANEMOI_GEMM_TIMELINE
    ANEMOI_B200_LIB=anemoi_core_b200/lib/variants/gemm_timeline.so python profiles/gemm_timeline.py
Prints, per GEMM shape, the cycles from kernel entry to each milestone (SM clock of CTA 0) and the CUDA-event time of the launch."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from anemoi_core_b200 import _lib  # noqa: E402
from anemoi_core_b200 import ops  # noqa: E402

NAMES = {0: "entry", 1: "setup done (barriers, TMEM, cluster sync)", 2: "griddepcontrol.wait returned", 3: "producer: first TMA issued", 4: "MMA: first stage landed",
         5: "MMA: last stage of tile 0 landed", 6: "epilogue: accumulator of tile 0 complete", 16: "epi round 0: residual landed", 7: "epi round 0: store issued",
         17: "epi round 1: residual landed", 8: "epi round 1: store issued", 18: "epi round 2: residual landed", 9: "epi round 2: store issued",
         19: "epi round 3: residual landed", 10: "epi round 3: store issued", 11: "epilogue: all stores complete", 12: "final cluster sync passed"}  # fmt: skip
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
lib = _lib.load()
lib.anemoi_b200_debug_timeline.argtypes = [ctypes.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, gelu, res in (("mlp2+res 1 wave", 9472, 512, 2048, False, True), ("mlp2+res 5 waves", 40962, 512, 2048, False, True),
                                 ("projection 1 wave", 9472, 512, 704, False, True), ("mlp1+gelu 4 waves", 9472, 2048, 512, True, False),
                                 ("qkv 1 tile per pair", 2048, 2240, 512, False, False)):  # fmt: skip
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) / K**0.5).to(torch.bfloat16).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    r = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev) if res else None
    o = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.linear(a, w, b, gelu=gelu, residual=r, out=o)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.linear(a, w, b, gelu=gelu, residual=r, out=o)
    e1.record()
    torch.cuda.synchronize()
    ts = (ctypes.c_longlong * 32)()
    assert lib.anemoi_b200_debug_timeline(ts) == 0
    t0 = ts[0]
    clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 0
    rec = {"gemm": name, "event_us": round(e0.elapsed_time(e1) * 1e3, 1), "sm_mhz_now": clk,
           "cycles_from_entry": {NAMES[i]: int(ts[i] - t0) for i in sorted(NAMES, key=lambda k: ts[k]) if ts[i] >= t0 and ts[i] != 0}}
    print(json.dumps(rec))
