"""The oracle restatement (oracle/restatement.py) must reproduce the golden vectors generated from the
UNMODIFIED reference modules (oracle/gen_golden.py).  fp32 CPU; tolerance 1e-5 abs on O(1) values (both
sides are torch fp32, only summation order in scatter/index_add may differ); integer path bit-exact."""
import pytest
import torch

from oracle import restatement as R

TOL = dict(atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["gnn_processor_small", "gnn_processor_cfg1"])
def test_gnn_processor(golden, name):
    g = golden(name)
    y = R.gnn_processor({"proc." + k[5:] if k.startswith("proc.") else k: v for k, v in g["sd"].items()}, g["x"], g["edge_attr"],
                        g["edge_index"], g["cfg"]["num_layers"])  # fmt: skip
    torch.testing.assert_close(y, g["y"], **TOL)


@pytest.mark.parametrize("name", ["gt_processor_small", "gt_processor_qknorm", "gt_processor_unsorted"])
def test_gt_processor(golden, name):
    g = golden(name)
    ei, ea = g["edge_index"], g["edge_attr"]
    if not g["sorted"]:
        ei, perm = R.sort_edge_index_by_dst(ei)
        ea = ea[perm]
    y = R.gt_processor(g["sd"], g["x"], ea, ei, g["cfg"]["num_layers"], g["cfg"]["num_heads"])
    torch.testing.assert_close(y, g["y"], **TOL)


def test_gnn_mappers(golden):
    g = golden("gnn_forward_mapper")
    ys, yd = R.gnn_forward_mapper(g["sd"], g["x_src"], g["x_dst"], g["edge_attr"], g["edge_index"])
    torch.testing.assert_close(ys, g["y_src"], **TOL)
    torch.testing.assert_close(yd, g["y_dst"], **TOL)
    g = golden("gnn_backward_mapper")
    y = R.gnn_backward_mapper(g["sd"], g["x_src"], g["x_dst"], g["edge_attr"], g["edge_index"])
    torch.testing.assert_close(y, g["y"], **TOL)


def test_gt_mappers(golden):
    for name in ("gt_forward_mapper_chunks1", "gt_forward_mapper_chunks4"):
        g = golden(name)
        _, yd = R.gt_forward_mapper(g["sd"], g["x_src"], g["x_dst"], g["edge_attr"], g["edge_index"], g["cfg"]["num_heads"])
        torch.testing.assert_close(yd, g["y_dst"], **TOL)
    g = golden("gt_backward_mapper")
    y = R.gt_backward_mapper(g["sd"], g["x_src"], g["x_dst"], g["edge_attr"], g["edge_index"], g["cfg"]["num_heads"])
    torch.testing.assert_close(y, g["y"], **TOL)


def test_gt_conv_both_statements(golden):
    """PyG-style softmax (conv.py:103-147) and the Triton kernel's online softmax (triton/gt.py:81-179)
    agree with the reference output at the reference's own bar (atol 1e-4, test_triton_gt.py:135-136)."""
    for c in golden("gt_conv")["cases"]:
        n_dst = c["q"].shape[0]
        out = R.gt_attention(c["q"], c["k"], c["v"], c["e"], c["edge_index"], n_dst)
        torch.testing.assert_close(out, c["out"], atol=1e-5, rtol=1e-5)
        (row, colptr), _, _ = R.edge_index_to_csc(c["edge_index"], (c["k"].shape[0], n_dst), True)
        out2 = R.gt_attention_online(c["q"], c["k"], c["v"], c["e"], row, colptr)
        torch.testing.assert_close(out2, c["out"], atol=1e-4, rtol=0)
        assert torch.all(out2[-2:] == 0)  # zero in-degree rows


def test_integer_path_bit_exact(golden):
    for c in golden("integer_path")["cases"]:
        ei, nn = c["edge_index"], c["num_nodes"]
        s, perm = R.sort_edge_index_by_dst(ei)
        assert torch.equal(s, c["sorted"]) and torch.equal(perm, c["perm"])
        (row, colptr), perm2, (rowptr, eids, edst) = R.edge_index_to_csc(ei, nn, False)
        for a, b in ((row, c["row"]), (colptr, c["colptr"]), (rowptr, c["rowptr"]), (eids, c["edge_ids"]), (edst, c["edge_dst"])):
            assert a.dtype == b.dtype and torch.equal(a, b)
        for parts, p in c["partitions"].items():
            dst_splits, edge_splits = R.build_graph_partition(s, parts, nn)
            assert dst_splits == p["dst_splits"] and edge_splits == p["edge_splits"]
            for cid, m in enumerate(p["chunks"]):
                (d0, d1), (e0, e1), connected, ei_c = R.materialise_chunk(dst_splits, edge_splits, cid, nn[0], s)
                assert torch.equal(torch.arange(d0, d1), m["dst_ids"])
                assert torch.equal(torch.arange(e0, e1), m["edge_ids"])
                assert torch.equal(connected, m["src_ids"])
                assert torch.equal(ei_c, m["edge_index"])
