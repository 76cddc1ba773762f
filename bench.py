"""bench.py — forward ms/step of the graph message-passing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|small]

Workload (default cfg2 = BASELINE.json configs[1]): O96 grid (40 320 points) -> icosahedral level-6 multi-scale mesh
(40 962 nodes, 327 600 edges), GraphTransformer encoder + 16 x 512 processor (16 heads) + decoder, bf16 autocast,
batch 1, synthetic inputs N(0,1), PyTorch-default random-init weights (seed 1234).  One "step" = encoder -> processor ->
latent skip -> decoder (models/encoder_processor_decoder.py:260-324).

Printed JSON (one line, rank 0):
  value / ms_per_step : device-resident step (inputs already in HBM), whole step replayed as one CUDA graph, CUDA events
                        per step, L2 flushed between timed steps, max over ranks.
  e2e                 : same step through the public module API with HOST (pinned) input buffers: H2D of the inputs,
                        the step, D2H of the output, all inside the timed region.
  roofline            : dominant kernel class measured live with CUDA events around every C-ABI launch (eager pass).
  cpu_baseline        : the UNMODIFIED reference modules (baseline/_ref + oracle/standins; the oracle port when they are absent) on the host
                        cores, fp32, rank 0, N=1.
  reference_gpu       : context, N=1: the same unmodified reference modules run on THIS GPU through their own code path (PyTorch + the
                        reference's Triton attention backend, bf16 autocast, eager) - ms/step and the difference to our output.
  --impl reference    : the reference's CPU path (the reference arm of the contract), whole steps.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (grid, mesh level, kind, channels, layers, heads, in_grid, in_mesh, out_grid)
    "cfg2": dict(grid="o96", mesh_level=6, kind="graphtransformer", C=512, layers=16, heads=16, in_grid=212, in_mesh=12, out_grid=88,
                 desc="O96 grid (40320 pts) -> ico-6 multi-scale mesh (40962 nodes, 327600 edges), GraphTransformer enc + 16x512 proc (16 heads) + dec"),
    "cfg3": dict(grid="n320", mesh_level=6, kind="gnn", C=1024, layers=16, heads=16, in_grid=212, in_mesh=12, out_grid=88,
                 desc="N320 grid (542080 pts) -> ico-6 mesh, GNN enc + 16x1024 proc + dec"),
    "cfg4": dict(grid="o1280", mesh_level=7, kind="graphtransformer", C=1024, layers=16, heads=16, in_grid=212, in_mesh=12, out_grid=88,
                 desc="O1280 grid (6599680 pts) -> ico-7 multi-scale mesh (163842 nodes, 1310640 edges), GraphTransformer enc + 16x1024 proc (16 heads) + dec; "
                      "BASELINE.json configs[3], meant for --gpus 8 (grid / mesh rows and edges sharded)"),
    "cfg5": dict(grid="o96", mesh_level=6, kind="graphtransformer", C=512, layers=16, heads=16, in_grid=212, in_mesh=12, out_grid=88, rollout=40,
                 desc="cfg2 model, 40-step autoregressive rollout (prognostic outputs written back into the newest input time slot, forcings/statics held), per-step latency"),
    "small": dict(grid="o32", mesh_level=4, kind="graphtransformer", C=512, layers=4, heads=16, in_grid=212, in_mesh=12, out_grid=88,
                  desc="O32 grid -> ico-4 mesh, GraphTransformer 4x512 (smoke-size)"),
}  # fmt: skip


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"], "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "tensor_burst": 1590.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, device_index: int):
        self.idx, self.proc, self.lines = device_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(w, gr, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(gr["n_grid"], w["in_grid"], generator=g), torch.randn(gr["n_mesh"], w["in_mesh"], generator=g)


def build_model(w, gr):
    from anemoi_core_b200.model import EncProcDec

    torch.manual_seed(1234)
    return EncProcDec(w["kind"], in_grid=w["in_grid"], in_mesh=w["in_mesh"], out_grid=w["out_grid"], num_channels=w["C"], num_layers=w["layers"],
                      edge_dim=gr["edge_dim"], num_heads=w["heads"]).eval()  # fmt: skip


def state_dicts(model):
    return {k: {n: p.detach().clone() for n, p in getattr(model, k).state_dict().items()} for k in ("encoder", "processor", "decoder")}


# ----------------------------------------------------------------------------------------------------------------
# CPU reference leg (oracle port)
# ----------------------------------------------------------------------------------------------------------------
def _ref_kwargs(w, gr):
    return dict(in_grid=w["in_grid"], in_mesh=w["in_mesh"], out_grid=w["out_grid"], num_channels=w["C"], num_layers=w["layers"],
                edge_dim=gr["edge_dim"], num_heads=w["heads"])  # fmt: skip


class CpuReference:
    """The reference's CPU forward step for a workload.  kind "reference": the UNMODIFIED reference modules (oracle/reference_step.py:
    /root/reference or baseline/_ref + oracle/standins); kind "port": oracle/restatement.py when the reference is not on this box.
    ``sample_layers`` < layers times encoder + that many processor layers + decoder and scales the processor linearly (stated in
    ``sample``); None runs the whole step."""

    def __init__(self, w, gr, sds, sample_layers=None):
        from oracle import reference_step as RS

        self.w, self.gr, self.sds = w, gr, sds
        self.sample_layers = None if (sample_layers is None or sample_layers >= w["layers"]) else sample_layers
        self.kind, self.ref = "port", None
        if RS.reference_root() is not None:
            try:
                self.ref = RS.ReferenceStep(w["kind"], state_dicts=sds, max_layers=self.sample_layers, **_ref_kwargs(w, gr))
                self.kind = "reference"
            except Exception as e:  # noqa: BLE001 - fall back to the port, and say why
                print(f"[bench] reference modules unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)

    def step(self):
        """(t_encoder, t_processor_sample, t_decoder) seconds of one step."""
        w, gr, sds = self.w, self.gr, self.sds
        if self.ref is not None:
            t = []
            self.ref(self.x_grid, self.x_mesh, gr, t)
            return t[0]
        from oracle import restatement as R

        H, L = w["heads"], w["layers"]
        with torch.no_grad():
            t0 = time.perf_counter()
            if w["kind"] == "graphtransformer":
                _, lat = R.gt_forward_mapper(sds["encoder"], self.x_grid, self.x_mesh, gr["enc_attr"], gr["enc_index"], H)
                t1 = time.perf_counter()
                proc = R.gt_processor(sds["processor"], lat, gr["proc_attr"], gr["proc_index"], L, H, max_layers=self.sample_layers)
                t2 = time.perf_counter()
                R.gt_backward_mapper(sds["decoder"], proc + lat, self.x_grid, gr["dec_attr"], gr["dec_index"], H)
            else:
                src_emb, lat = R.gnn_forward_mapper(sds["encoder"], self.x_grid, self.x_mesh, gr["enc_attr"], gr["enc_index"])
                t1 = time.perf_counter()
                proc = R.gnn_processor(sds["processor"], lat, gr["proc_attr"], gr["proc_index"], L, max_layers=self.sample_layers)
                t2 = time.perf_counter()
                R.gnn_backward_mapper(sds["decoder"], proc + lat, src_emb, gr["dec_attr"], gr["dec_index"])
            t3 = time.perf_counter()
        return t1 - t0, t2 - t1, t3 - t2

    def ms(self, x_grid, x_mesh):
        self.x_grid, self.x_mesh = x_grid, x_mesh
        te, tl, td = self.step()
        scale = 1.0 if self.sample_layers is None else self.w["layers"] / self.sample_layers
        return (te + td + tl * scale) * 1e3, (te, tl, td)

    def describe(self, parts) -> str:
        te, tl, td = parts
        what = ("the unmodified reference modules (anemoi.models.layers.{mapper,processor}, pyg attention backend) via oracle/reference_step.py"
                if self.kind == "reference" else "oracle/restatement.py (port of the reference PyTorch path)")
        if self.sample_layers is None:
            return f"{what}, fp32 on CPU, ONE WHOLE step: encoder {te:.2f}s + all {self.w['layers']} processor layers {tl:.2f}s + decoder {td:.2f}s"
        return (f"{what}, fp32 on CPU: encoder {te:.2f}s + {self.sample_layers} of {self.w['layers']} processor layers {tl:.2f}s + decoder "
                f"{td:.2f}s; processor scaled x{self.w['layers'] / self.sample_layers:g} (layers are identical)")


def cpu_baseline(w, gr, sds, x_grid, x_mesh):
    """Bounded CPU sample beside the GPU number (rank 0, N = 1): one warm-up + one timed WHOLE step when a step is ~10 s (cfg2),
    encoder + 2 layers + decoder scaled for the 16 x 1024 workloads."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = CpuReference(w, gr, sds, sample_layers=None if w["C"] <= 512 else 2)
    ref.ms(x_grid, x_mesh)  # warm-up
    ms, parts = ref.ms(x_grid, x_mesh)
    return {"value": ms, "unit": "ms/step", "cores": torch.get_num_threads(), "kind": ref.kind, "sample": ref.describe(parts)}


def run_reference(args, w):
    """--impl reference: the reference's own CPU implementation of the step on all host threads (unmodified reference modules when
    baseline/_ref or /root/reference is present, else the oracle port), WHOLE steps (no extrapolation for cfg2), as many of the K
    requested as fit a ~150 s budget; ``steps`` reports how many were timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph(w["grid"], w["mesh_level"])
    model = build_model(w, gr)
    sds = state_dicts(model)
    del model
    x_grid, x_mesh = make_inputs(w, gr)
    torch.set_num_threads(os.cpu_count() or 1)
    ref = CpuReference(w, gr, sds, sample_layers=None if w["C"] <= 512 else 2)
    n_warm = min(max(args.warmup, 1), 2)
    for _ in range(n_warm):
        ref.ms(x_grid, x_mesh)
    vals, parts = [], None
    t_budget = time.perf_counter()
    for _ in range(max(args.steps, 1)):
        ms, parts = ref.ms(x_grid, x_mesh)
        vals.append(ms)
        if time.perf_counter() - t_budget > 150.0:  # keep the arm within a few minutes whatever K is
            break
    ms = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": "forward ms/step", "value": ms, "unit": "ms/step", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": n_warm, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": f"{args.workload}: {w['desc']}", "batch": 1, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": ms, "unit": "ms/step", "cores": torch.get_num_threads(), "kind": ref.kind,
                         "sample": ref.describe(parts) + f"; mean of {len(vals)} timed steps after {n_warm} warm-up"},
        "e2e": {"value": ms, "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def single_gpu_rows(model, w, gd, x_full, x_grid, x_mesh, grid_shards, dev, n_out_rows):
    """Rank 0's parity reference at N > 1: the SAME step computed on this one GPU without a group, restricted to the output rows the
    sharded run returns on this rank.  Sharded inputs (GraphTransformer): encoder and processor over the whole graph, decoder over rank 0's
    grid rows only (the edges into them, relabelled) - identical to the whole-graph decoder on those rows, and it keeps the 6.6 M-row cfg4
    decoder off a single GPU.  Replicated inputs: the plain single-GPU forward."""
    from anemoi_core_b200 import ops
    from anemoi_core_b200.distributed.shapes import BipartiteGraphShardInfo
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        if x_full is None:
            return model(x_grid, x_mesh, gd)[:n_out_rows]
        xg, xm = x_full[0].to(dev), x_full[1].to(dev)
        bi = BipartiteGraphShardInfo()
        _, lat = model.encoder((xg, xm), 1, bi, gd["enc_attr"], gd["enc_index"])
        y = model.processor(lat, 1, GraphShardInfo(nodes=[lat.shape[0]]), gd["proc_attr"], gd["proc_index"])
        y = ops.add(y, lat)
        n0 = grid_shards[0]
        e1 = int(torch.searchsorted(gd["dec_index"][1].contiguous(), torch.tensor([n0], device=dev)).item())  # dst-sorted: rows [0, n0) own edges [0, e1)
        return model.decoder((y, xg[:n0]), 1, bi, gd["dec_attr"][:e1], gd["dec_index"][:, :e1].contiguous())


# ----------------------------------------------------------------------------------------------------------------
def reference_gpu(workload: str, w, timeout_s: int = 240):
    """CONTEXT number, N = 1 only: the UNMODIFIED reference modules (baseline/_ref + oracle/standins) run on the same GPU through their own code
    path - PyTorch / cuBLAS and the reference's Triton attention backend (its fastest) - under bf16 autocast, eager, CUDA events, L2 flushed
    (profiles/bench_reference_gpu.py --bench-leg, in a SUBPROCESS with a timeout: whatever happens there cannot take the bench line down).
    The reference arm of the contract stays the CPU path (--impl reference); this says what the reference's own GPU path does on this box."""
    try:
        from oracle import reference_step as RS

        if RS.reference_root() is None:
            return {"unavailable": "baseline/_ref not on this box"}
        cmd = [sys.executable, os.path.join(ROOT, "profiles", "bench_reference_gpu.py"), "--workload", workload, "--steps", "5", "--bench-leg"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": f"rc {r.returncode}: {(r.stderr or r.stdout)[-200:]}"}
        d = json.loads(lines[-1])
        backend = "triton" if w["kind"] == "graphtransformer" else "pyg"
        ref = d.get(f"reference_{backend}_bf16")
        if not isinstance(ref, dict):
            return {"unavailable": str(d.get(f"reference_{backend}", "no reference line"))[:200]}
        return {"value": ref["ms_per_step_eager"], "unit": "ms/step", "launch": "eager", "attention_backend": backend, "dtype": "bf16 autocast",
                "kind": "unmodified reference modules on the same GPU (PyTorch + its Triton kernel)", "steps": 5,
                "ours_eager_ms_same_process": d.get("ours_eager_bf16_ms"), "rel_l2_ours_vs_reference": ref["rel_l2_ours_vs_reference"]}
    except Exception as e:  # noqa: BLE001 - context only: never fail the bench line
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))  # cfg2 = BASELINE.json configs[1] (the metric's config)
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the context run of the unmodified reference on the GPU (N = 1)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    from anemoi_core_b200 import ops
    from anemoi_core_b200.distributed.balanced_partition import get_balanced_partition_sizes
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph(w["grid"], w["mesh_level"])
    model = build_model(w, gr)
    sds = state_dicts(model) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    model = model.to(dev)
    from anemoi_core_b200.layers._functional import freeze_packed_weights

    gd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in gr.items()}
    x_grid_h, x_mesh_h = make_inputs(w, gr)
    mesh_shards = get_balanced_partition_sizes(gr["n_mesh"], world) if world > 1 else None
    grid_shards = get_balanced_partition_sizes(gr["n_grid"], world) if world > 1 else None
    # N > 1, GraphTransformer: every rank holds, uploads and returns ITS rows only (the reference's in_out_sharded mode,
    # encoder_processor_decoder.py:176-183); what crosses NVLink is the halo of each exchange.  GNN: replicated inputs (its mappers shard inside).
    sharded_io = world > 1 and w["kind"] == "graphtransformer"
    x_full = None
    if sharded_io:
        g0, m0 = sum(grid_shards[:rank]), sum(mesh_shards[:rank])
        if rank == 0:
            x_full = (x_grid_h, x_mesh_h)  # kept on the host for the single-GPU parity check below
        x_grid_h, x_mesh_h = x_grid_h[g0 : g0 + grid_shards[rank]].clone(), x_mesh_h[m0 : m0 + mesh_shards[rank]].clone()
        fwd_kw = dict(model_comm_group=group, mesh_shards=mesh_shards, grid_shards=grid_shards, keep_output_sharded=True, inputs_sharded=True)
    elif world > 1:
        fwd_kw = dict(model_comm_group=group, mesh_shards=mesh_shards, grid_shards=grid_shards)
    else:
        fwd_kw = {}
    x_grid_h, x_mesh_h = x_grid_h.pin_memory(), x_mesh_h.pin_memory()
    x_grid, x_mesh = x_grid_h.to(dev), x_mesh_h.to(dev)

    def step(xg, xm):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(xg, xm, gd, **fwd_kw)

    # ---- warm-up (also builds CSR plans, packed weights, TMA descriptors) and launch count per step --------------------
    for _ in range(2):
        out = step(x_grid, x_mesh)
    torch.cuda.synchronize()
    freeze_packed_weights(model)  # inference: weights are final, skip the per-call version checks on the host
    n0 = ops.LAUNCHES
    out = step(x_grid, x_mesh)
    launches_per_step = ops.LAUNCHES - n0
    torch.cuda.synchronize()

    # N = 1: the whole step is one CUDA graph.  N > 1, GraphTransformer: every exchange is a peer-memory kernel (csrc/peer.cu), so the
    # sharded step is ONE graph per rank as well.  Where NCCL collectives remain (GNN all-gathers, ANEMOI_B200_PEER=0) the compute between
    # two collectives is captured and the collectives run eagerly between the segment replays (SegmentedCapture).
    use_graph = not args.no_graph
    whole_graph = world == 1
    if world > 1 and sharded_io:
        from anemoi_core_b200.distributed import peer

        whole_graph = peer.available(group, dev)
    if use_graph:
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                if whole_graph:
                    replay = model.capture(x_grid, x_mesh, gd, **fwd_kw)
                else:
                    replay = model.capture_segmented(x_grid, x_mesh, gd, **fwd_kw)
        except Exception as e:  # noqa: BLE001  (e.g. a collective that cannot be captured): fall back to eager launches, and say so
            if world == 1:
                raise
            print(f"[bench] CUDA-graph capture failed on rank {rank} ({type(e).__name__}: {e}); using eager launches", file=sys.stderr)
            use_graph = False
    if world > 1:  # all ranks must agree
        flag = torch.tensor([1 if use_graph else 0], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        use_graph = bool(flag.item())
    if use_graph:
        run = lambda: replay()  # noqa: E731
    else:
        run = lambda: step(x_grid, x_mesh)  # noqa: E731

    # ---- the timed path must be the parity-tested path: graph replay == eager launches of the same step, and (N > 1) the sharded step ==
    # the single-GPU step computed by rank 0 on the whole graph (the path tests/test_gpu_parity.py pins against the oracle) ---------------
    def rel_err(a, b):
        a, b = a.float(), b.float()
        return {"max_rel": ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item(), "rel_l2": ((a - b).norm() / b.norm().clamp_min(1e-30)).item()}

    eager_out = step(x_grid, x_mesh).clone()
    checks = {}
    if use_graph:
        checks["replay_vs_eager"] = rel_err(run().clone(), eager_out)
        if checks["replay_vs_eager"]["max_rel"] > 1e-6:
            raise SystemExit(f"[bench] CUDA-graph replay differs from the eager step: {checks['replay_vs_eager']}")
    if world > 1:
        if rank == 0:
            checks["sharded_vs_single_gpu"] = rel_err(eager_out, single_gpu_rows(model, w, gd, x_full, x_grid, x_mesh, grid_shards, dev, out.shape[0]))
        bad = torch.tensor([1 if (rank == 0 and checks["sharded_vs_single_gpu"]["rel_l2"] > 5e-3) else 0], device=dev)
        torch.distributed.all_reduce(bad, op=torch.distributed.ReduceOp.MAX)
        if bad.item():
            raise SystemExit(f"[bench] sharded step differs from the single-GPU step: {checks.get('sharded_vs_single_gpu')}")
    del eager_out

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = [a.elapsed_time(b) for a, b in evs]
        return sum(ms) / len(ms), wall

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_dev, wall_dev = timed(run, args.steps, args.warmup)

    # ---- e2e: host buffers in, host buffer out ----------------------------------------------------------------------------
    out_h = torch.empty(out.shape, dtype=out.dtype).pin_memory()

    if use_graph:

        def e2e_step():
            o = replay(x_grid_h, x_mesh_h)  # H2D into the graph's static inputs, replay
            out_h.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    else:

        def e2e_step():
            o = step(x_grid_h.to(dev, non_blocking=True), x_mesh_h.to(dev, non_blocking=True))
            out_h.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    ms_e2e, _ = timed(e2e_step, args.steps, 3)

    # ---- cfg5: autoregressive rollout (training/tasks/forecaster.py:155-205 semantics): slot t <- slot t+1, newest slot <- prediction ----
    rollout = None
    if w.get("rollout"):
        nv = 100  # variables per time slot; the model predicts the first out_grid of them (prognostic), the rest are forcings
        xg = x_grid.clone()
        lat = []
        for i in range(w["rollout"]):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            y = replay(xg) if use_graph else step(xg, x_mesh)
            xg[:, :nv].copy_(xg[:, nv : 2 * nv])  # shift the time window
            xg[:, nv : nv + w["out_grid"]].copy_(y)  # newest slot <- prognostic prediction (forcings and statics stay)
            b.record()
            torch.cuda.synchronize()
            lat.append(a.elapsed_time(b))
        steady = sorted(lat[1:])
        rollout = {"steps": w["rollout"], "mean_ms": sum(steady) / len(steady), "p50_ms": steady[len(steady) // 2],
                   "p99_ms": steady[min(len(steady) - 1, int(0.99 * len(steady)))], "first_ms": lat[0]}
    clk = clocks.stop() if rank == 0 else None

    io_bytes = [x_grid_h.numel() * 4 + x_mesh_h.numel() * 4, out_h.numel() * out_h.element_size()]
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_dev, ms_e2e = t.tolist()
        t = torch.tensor(io_bytes, dtype=torch.int64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        io_bytes = t.tolist()

    # ---- roofline leg: CUDA events around every C-ABI launch, eager, over 3 steps -------------------------------------------
    pk = peaks()
    for _ in range(2):
        step(x_grid, x_mesh)
    torch.cuda.synchronize()
    ops.start_timing()
    n_prof = 3
    for _ in range(n_prof):
        flush.zero_()
        step(x_grid, x_mesh)
    torch.cuda.synchronize()
    rec = ops.stop_timing()
    agg = {}
    for name, a, b, fl, by in rec:
        d = agg.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += a.elapsed_time(b)
        d["flops"] += fl
        d["bytes"] += by
        d["launches"] += 1
    total_ms = sum(d["ms"] for d in agg.values()) or 1.0
    kernels = {}
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
        gb = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        kernels[name] = {"share": round(d["ms"] / total_ms, 4), "us_per_launch": round(1e3 * d["ms"] / d["launches"], 2),
                         "launches_per_step": d["launches"] // n_prof, "tflops": round(tf, 2), "gbs": round(gb, 1),
                         "frac_tensor": round(tf / pk["tensor"], 4), "frac_hbm": round(gb / pk["hbm"], 4)}  # fmt: skip
    top = next(iter(kernels))
    topd = agg[top]
    # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed `ncu --set full` capture of THIS workload
    # (profiles/r2/traffic_<workload>.json, written by profiles/ncu_traffic.py); null when no matching capture exists
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2", f"traffic_{args.workload}.json")
    if world == 1 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("kernel_class") == top:
            traffic, traffic_src = tj["avg_dram_bytes_per_launch"], os.path.relpath(tpath, ROOT)
    if world > 1:
        # per-launch event timing of eager launches includes host gaps at this size; the honest device-side figure at N > 1 is the whole
        # replayed step: this rank's algorithmic GEMM flops / the step time, against the sustained tensor peak
        fl = agg.get("linear_tcgen05", {"flops": 0.0})["flops"] / n_prof
        ach = fl / (ms_dev * 1e-3) / 1e12
        roofline = {"kernel": "whole step (graph replay), GEMM flops of this rank's shard", "bound": "tensor", "achieved": ach, "peak": pk["tensor"],
                    "unit": "TFLOP/s", "frac": ach / pk["tensor"], "traffic": None, "peak_source": pk["src"] + " sustained bf16"}  # fmt: skip
    elif top == "linear_tcgen05":
        ach = topd["flops"] / (topd["ms"] * 1e-3) / 1e12
        roofline = {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": ach, "peak": pk["tensor"], "unit": "TFLOP/s",
                    "frac": ach / pk["tensor"], "traffic": traffic, "traffic_unit": "bytes per launch (DRAM read+write, ncu)", "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": topd["bytes"] / topd["launches"], "peak_source": pk["src"] + " sustained bf16 (kernel timed inside a long step)",
                    "share_of_step": kernels[top]["share"]}  # fmt: skip
    else:
        ach = topd["bytes"] / (topd["ms"] * 1e-3) / 1e9
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic,
                    "traffic_source": traffic_src, "algorithmic_bytes_per_launch": topd["bytes"] / topd["launches"],
                    "peak_source": pk["src"], "share_of_step": kernels[top]["share"]}  # fmt: skip

    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    line = {
        "metric": "forward ms/step", "value": ms_dev, "unit": "ms/step", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "batch": 1, "precision": "bf16 autocast, fp32 accumulate",
                   "launch": ("cuda-graph replay (one graph per rank)" if whole_graph else "cuda-graph segments + eager NCCL collectives") if use_graph else "eager",
                   "l2": "flushed between timed steps (256 MiB memset outside the event pair)",
                   "parallelism": "single GPU" if world == 1 else (
                       f"grid and mesh rows, edges and every stage dst-range sharded over {world} GPUs; inputs / outputs stay sharded; per exchange the k|v halo rows "
                       f"are written into the peers' tables over NVLink ({'peer-memory kernels inside the graph' if whole_graph else 'NCCL all-to-all'})")},
        "e2e": {"value": ms_e2e, "unit": "ms/step", "h2d_bytes_per_step": io_bytes[0], "d2h_bytes_per_step": io_bytes[1],
                "note": "bytes summed over ranks; every rank copies its own rows" if sharded_io else "bytes summed over ranks"},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": clk,
        "roofline": roofline,
        "kernels": kernels,
        "parity": checks,
        "wall_s_timed_region": wall_dev,
    }  # fmt: skip
    if rollout is not None:
        line["rollout"] = rollout
    if world == 1 and not args.no_reference_gpu and w["C"] <= 512 and "rollout" not in w:
        line["reference_gpu"] = reference_gpu(args.workload, w)
    if sds is not None:
        line["cpu_baseline"] = cpu_baseline(w, gr, sds, x_grid_h, x_mesh_h)
    else:
        line["cpu_baseline"] = {"value": None, "unit": "ms/step", "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
