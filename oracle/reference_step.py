"""oracle/reference_step.py — the UNMODIFIED reference modules wired into the benchmark's forward step, on CPU.

*** TEST / MEASUREMENT INFRASTRUCTURE (bench.py --impl reference and its cpu_baseline leg; tests).  Never imported by anemoi_core_b200. ***

The reference hot path (``anemoi.models.layers.{mapper,processor}``) is imported from ``/root/reference/models/src`` when that is
mounted (build container) or from ``baseline/_ref`` (the `pip install --no-deps --target baseline/_ref` copy that travels to the GPU
box), with ``oracle/standins`` supplying torch_geometric / hydra / anemoi.utils.  The step is the call
sequence of ``AnemoiModelEncProcDec.forward`` (models/encoder_processor_decoder.py:260-324): encoder mapper -> processor -> latent
skip -> decoder mapper, fp32, ``graph_attention_backend="pyg"`` (the reference's own CPU path; its Triton backend needs CUDA and is
what ``profiles/bench_reference_gpu.py`` selects for the GPU head-to-head).
"""

from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def reference_root():
    """Directory holding the reference's ``anemoi/models`` package, or None."""
    for cand in ("/root/reference/models/src", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "anemoi", "models", "layers")):
            return cand
    return None


def _import_reference():
    root = reference_root()
    if root is None:
        raise ImportError("reference not found (neither /root/reference/models/src nor baseline/_ref)")
    for p in (root, os.path.join(HERE, "standins")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from anemoi.models.distributed.shapes import BipartiteGraphShardInfo
    from anemoi.models.distributed.shapes import GraphShardInfo
    from anemoi.models.layers import mapper as M
    from anemoi.models.layers import processor as P

    return M, P, GraphShardInfo, BipartiteGraphShardInfo


class ReferenceStep:
    """Reference encoder / processor / decoder for a bench workload, parameters copied from our modules' ``state_dict`` (same keys)."""

    def __init__(self, kind: str, *, in_grid: int, in_mesh: int, out_grid: int, num_channels: int, num_layers: int, edge_dim: int,
                 num_heads: int, state_dicts: dict, max_layers=None, attention_backend: str = "pyg") -> None:  # fmt: skip
        import torch

        M, P, self.GSI, self.BSI = _import_reference()
        self.kind, C = kind, num_channels
        self.root = reference_root()
        layers = num_layers if max_layers is None else min(num_layers, max_layers)
        if kind == "graphtransformer":
            common = dict(num_heads=num_heads, mlp_hidden_ratio=4.0, edge_dim=edge_dim, num_chunks=1, layer_kernels=None, graph_attention_backend=attention_backend)
            self.encoder = M.GraphTransformerForwardMapper(in_channels_src=in_grid, in_channels_dst=in_mesh, hidden_dim=C, **common)
            self.processor = P.GraphTransformerProcessor(num_layers=layers, num_channels=C, **common)
            self.decoder = M.GraphTransformerBackwardMapper(in_channels_src=C, in_channels_dst=in_grid, hidden_dim=C, out_channels_dst=out_grid, **common)
        else:
            common = dict(mlp_extra_layers=0, edge_dim=edge_dim, num_chunks=1, layer_kernels=None)
            self.encoder = M.GNNForwardMapper(in_channels_src=in_grid, in_channels_dst=in_mesh, hidden_dim=C, **common)
            self.processor = P.GNNProcessor(num_layers=layers, num_channels=C, **common)
            self.decoder = M.GNNBackwardMapper(in_channels_src=C, in_channels_dst=C, hidden_dim=C, out_channels_dst=out_grid, **common)
        self.encoder.load_state_dict(state_dicts["encoder"], strict=True)
        proc_sd = state_dicts["processor"]
        if layers != num_layers:  # bounded sample: the first `layers` blocks
            proc_sd = {k: v for k, v in proc_sd.items() if not k.startswith("proc.") or int(k.split(".")[1]) < layers}
        self.processor.load_state_dict(proc_sd, strict=True)
        self.decoder.load_state_dict(state_dicts["decoder"], strict=True)
        for m in (self.encoder, self.processor, self.decoder):
            m.eval()
        self._torch = torch

    def to(self, device):
        for m in (self.encoder, self.processor, self.decoder):
            m.to(device)
        return self

    def __call__(self, x_grid, x_mesh, gr, times=None):
        """One forward step; ``times`` (a list) receives (t_encoder, t_processor, t_decoder) in seconds."""
        import time

        torch = self._torch
        bi = self.BSI(src_nodes=None, dst_nodes=None, edges=None)
        with torch.set_grad_enabled(not getattr(self, "no_grad", True)):
            t0 = time.perf_counter()
            x_data_latent, x_latent = self.encoder((x_grid, x_mesh), 1, bi, gr["enc_attr"], gr["enc_index"], None)
            t1 = time.perf_counter()
            x_proc = self.processor(x_latent, 1, self.GSI(nodes=[x_latent.shape[0]], edges=None), gr["proc_attr"], gr["proc_index"], None)
            t2 = time.perf_counter()
            x_proc = x_proc + x_latent
            out = self.decoder((x_proc, x_data_latent), 1, bi, gr["dec_attr"], gr["dec_index"], None)
            t3 = time.perf_counter()
        if times is not None:
            times.append((t1 - t0, t2 - t1, t3 - t2))
        return out
