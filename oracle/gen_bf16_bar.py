"""oracle/gen_bf16_bar.py — fixture for the bf16 parity criterion of SURVEY.md §8(d):

    err(ours under bf16 autocast, vs reference fp32)  <=  1.5 x err(REFERENCE under bf16 autocast, vs reference fp32)

*** TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference). ***     python oracle/gen_bf16_bar.py

The UNMODIFIED reference processors run twice on CPU — fp32, and under ``torch.autocast("cpu", dtype=torch.bfloat16)`` (Linear in
bf16 with fp32 accumulation, LayerNorm / softmax / scatter in the dtype autocast leaves them, i.e. the reference's own mixed-precision
path) — on the ico-4 multi-scale mesh of ``anemoi_core_b200.synthetic`` with seeded default-init weights.  The fixture stores the inputs'
seeds, a parameter checksum (so a drifting RNG is detected instead of silently comparing different models), the fp32 output and the
autocast output (bf16).  Weights are NOT stored (13 MB): the test re-creates them from the seed.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(1, "/root/reference/models/src")
sys.path.insert(2, ROOT)

import torch  # noqa: E402

from anemoi.models.distributed.shapes import GraphShardInfo  # noqa: E402
from anemoi.models.layers.processor import GNNProcessor  # noqa: E402
from anemoi.models.layers.processor import GraphTransformerProcessor  # noqa: E402
from anemoi_core_b200.synthetic import build_graph  # noqa: E402

CASES = {
    "gt": dict(seed=101, C=256, H=8, layers=4),
    "gnn": dict(seed=102, C=128, layers=3),
}


def make(kind, cfg, edge_dim, ref=True):
    torch.manual_seed(cfg["seed"])
    if kind == "gt":
        kw = dict(num_layers=cfg["layers"], num_channels=cfg["C"], num_chunks=1, num_heads=cfg["H"], mlp_hidden_ratio=4, edge_dim=edge_dim)
        if ref:
            kw.update(layer_kernels=None, graph_attention_backend="pyg")
        return kw
    kw = dict(num_channels=cfg["C"], num_layers=cfg["layers"], num_chunks=1, mlp_extra_layers=0, edge_dim=edge_dim)
    if ref:
        kw.update(layer_kernels=None)
    return kw


def checksum(m) -> float:
    return float(sum(p.detach().double().abs().sum() for p in m.parameters()))


@torch.no_grad()
def main():
    gr = build_graph("o32", mesh_level=4)
    n = gr["n_mesh"]
    out = {"kind": "bf16_bar", "graph": ("o32", 4), "cases": {}}
    for kind, cfg in CASES.items():
        cls = GraphTransformerProcessor if kind == "gt" else GNNProcessor
        m = cls(**make(kind, cfg, gr["edge_dim"])).eval()
        x = torch.randn(n, cfg["C"], generator=torch.Generator().manual_seed(cfg["seed"] + 1))
        si = GraphShardInfo(nodes=[n], edges=None)
        y32 = m(x, 1, si, gr["proc_attr"], gr["proc_index"], None)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yac = m(x, 1, si, gr["proc_attr"], gr["proc_index"], None)
        err = ((yac.float() - y32).norm() / y32.norm()).item()
        out["cases"][kind] = {"cfg": cfg, "param_checksum": checksum(m), "y32": y32.clone(), "y_autocast": yac.to(torch.bfloat16).clone(),
                              "ref_autocast_rel_l2": err}
        print(kind, tuple(y32.shape), "reference-under-autocast rel-L2 vs fp32:", err)
    torch.save(out, os.path.join(ROOT, "tests", "golden", "bf16_bar.pt"))


if __name__ == "__main__":
    main()
