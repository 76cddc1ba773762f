"""Parity bars the round-1 review asked for (-m gpu):

* the TIMED path is the tested path: ``EncProcDec.capture`` (CUDA-graph replay on static buffers, after ``freeze_packed_weights``) returns
  bit-for-bit what the eager forward returns, also for new inputs copied into the static buffers;
* SURVEY.md §8(d) bf16 criterion: err(ours under bf16 autocast vs reference fp32) <= 1.5 x err(REFERENCE under bf16 autocast vs reference
  fp32), on the fixtures of oracle/gen_bf16_bar.py (unmodified reference processors, fp32 and autocast outputs);
* cfg3 at full size (ico-6 mesh, GNN, C = 1024) on a 2-layer sample against the oracle: fp32 <= 1e-4, bf16 rel-L2 <= 8e-3.
"""
import pytest
import torch

from oracle import restatement as R

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("kind", ["graphtransformer", "gnn"])
def test_capture_replay_equals_eager(kind):
    from anemoi_core_b200.layers._functional import freeze_packed_weights
    from anemoi_core_b200.model import EncProcDec
    from anemoi_core_b200.synthetic import build_graph

    gr = build_graph("o32", 4)
    torch.manual_seed(7)
    m = EncProcDec(kind, in_grid=20, in_mesh=12, out_grid=9, num_channels=256, num_layers=3, edge_dim=gr["edge_dim"], num_heads=8).cuda().eval()
    gd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in gr.items()}
    g = torch.Generator().manual_seed(9)
    xg, xm = torch.randn(gr["n_grid"], 20, generator=g).cuda(), torch.randn(gr["n_mesh"], 12, generator=g).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        eager = m(xg, xm, gd).clone()
        freeze_packed_weights(m)
        frozen = m(xg, xm, gd).clone()
        replay = m.capture(xg, xm, gd)
        got = replay().clone()
        got2 = replay().clone()
        xg2 = xg * 0.5 + 1.0
        new_in = replay(xg2, xm).clone()
        eager2 = m(xg2, xm, gd)
    assert torch.equal(eager, frozen), "freeze_packed_weights changed the result"
    assert torch.equal(eager, got) and torch.equal(got, got2), "CUDA-graph replay differs from the eager forward"
    assert torch.equal(new_in, eager2), "replay with new inputs differs from the eager forward on them"
    freeze_packed_weights(m, False)


@pytest.mark.parametrize("kind", ["gt", "gnn"])
def test_bf16_error_within_1p5x_of_reference_autocast(golden, kind):
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.layers import GraphTransformerProcessor
    from anemoi_core_b200.synthetic import build_graph

    fx = golden("bf16_bar")
    c = fx["cases"][kind]
    cfg = c["cfg"]
    gr = build_graph(*fx["graph"])
    n = gr["n_mesh"]
    torch.manual_seed(cfg["seed"])  # same seeded default init as oracle/gen_bf16_bar.py
    if kind == "gt":
        m = GraphTransformerProcessor(num_layers=cfg["layers"], num_channels=cfg["C"], num_chunks=1, num_heads=cfg["H"], mlp_hidden_ratio=4, edge_dim=gr["edge_dim"])
    else:
        m = GNNProcessor(num_channels=cfg["C"], num_layers=cfg["layers"], num_chunks=1, mlp_extra_layers=0, edge_dim=gr["edge_dim"])
    checksum = float(sum(p.detach().double().abs().sum() for p in m.parameters()))
    assert abs(checksum - c["param_checksum"]) <= 1e-6 * c["param_checksum"], "seeded init differs from the fixture's (RNG drift): regenerate the fixture"
    x = torch.randn(n, cfg["C"], generator=torch.Generator().manual_seed(cfg["seed"] + 1))
    m = m.cuda().eval()
    from anemoi_core_b200.distributed.shapes import GraphShardInfo

    with torch.no_grad():
        y32 = m(x.cuda(), 1, GraphShardInfo(nodes=[n]), gr["proc_attr"].cuda(), gr["proc_index"].cuda())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(x.cuda(), 1, GraphShardInfo(nodes=[n]), gr["proc_attr"].cuda(), gr["proc_index"].cuda())
    assert rel_l2(y32, c["y32"]) <= 1e-4, "fp32 path vs the reference's fp32 output"
    ours, ref = rel_l2(y16, c["y32"]), rel_l2(c["y_autocast"], c["y32"])
    assert abs(ref - c["ref_autocast_rel_l2"]) <= 0.2 * c["ref_autocast_rel_l2"]  # (the stored autocast output is itself bf16-rounded)
    assert ours <= 1.5 * ref, f"bf16 error {ours:.3e} exceeds 1.5 x the reference-under-autocast error {ref:.3e}"
    assert ours <= 8e-3


def test_cfg3_full_size_two_layer_sample():
    """BASELINE cfg3 processor shape at FULL size (ico-6 multi-scale mesh: 40 962 nodes, 327 600 edges; GNN, C = 1024), 2 of the 16 layers."""
    from anemoi_core_b200.distributed.shapes import GraphShardInfo
    from anemoi_core_b200.layers import GNNProcessor
    from anemoi_core_b200.synthetic import icosphere_multiscale
    import numpy as np

    v, e = icosphere_multiscale(6)
    perm = np.argsort(e[1], kind="stable")
    ei = torch.from_numpy(np.stack([e[0][perm], e[1][perm]]).astype(np.int64))
    n, E, C, d_e = v.shape[0], ei.shape[1], 1024, 11
    assert (n, E) == (40962, 327600)
    g = torch.Generator().manual_seed(11)
    ea, x = torch.randn(E, d_e, generator=g), torch.randn(n, C, generator=g)
    torch.manual_seed(12)
    m = GNNProcessor(num_channels=C, num_layers=2, num_chunks=1, mlp_extra_layers=0, edge_dim=d_e).eval()
    sd = {k: p.clone() for k, p in m.state_dict().items()}
    with torch.no_grad():
        ref = R.gnn_processor(sd, x, ea, ei, 2)
    m = m.cuda()
    with torch.no_grad():
        y32 = m(x.cuda(), 1, GraphShardInfo(nodes=[n]), ea.cuda(), ei.cuda())
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16 = m(x.cuda(), 1, GraphShardInfo(nodes=[n]), ea.cuda(), ei.cuda())
    mx = ((y32.float().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert mx <= 1e-4 and rel_l2(y32, ref) <= 1e-4, f"cfg3 fp32: max-rel {mx:.3e}, rel-L2 {rel_l2(y32, ref):.3e}"
    assert rel_l2(y16, ref) <= 8e-3, f"cfg3 bf16 rel-L2 {rel_l2(y16, ref):.3e}"
