"""oracle/fuzz_integer_path.py — randomised comparison of the integer / index path with the UNMODIFIED reference functions.

*** TEST INFRASTRUCTURE (run by tests/test_host_logic.py in a subprocess, so that the reference's ``anemoi`` package never shares a process
with the overlay tests).  Never imported by anemoi_core_b200. ***

Reference side (imported from /root/reference/models/src or baseline/_ref with oracle/standins): ``distributed/khop_edges.py``
``sort_edge_index_by_dst`` (:37-40), ``is_edge_index_dst_sorted`` (:43-48), ``build_graph_partition`` (:154-189), ``GraphPartition.materialise``
(:78-132), ``distributed/balanced_partition.py:get_balanced_partition_sizes`` (:16-41).  Ours: ``anemoi_core_b200.distributed.khop_edges`` /
``balanced_partition``.  Bar: bit-exact ("integer, bit-exact").  The committed fixtures (tests/golden/integer_path.pt) pin four
seeded graphs; this walks a few hundred random ones, including the degenerate shapes: no edges into a whole part, every edge into one
destination, more parts than destinations, a single source, duplicate edges.
    python oracle/fuzz_integer_path.py [cases]   ->   one JSON line {"cases": n, "checks": m, "failures": [...]}"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.reference_step import reference_root  # noqa: E402


def main(n_cases: int) -> dict:
    root = reference_root()
    if root is None:
        return {"unavailable": "reference not found (neither /root/reference/models/src nor baseline/_ref)"}
    for p in (root, os.path.join(HERE, "standins")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from anemoi.models.distributed import balanced_partition as RB
    from anemoi.models.distributed import khop_edges as RK

    from anemoi_core_b200.distributed import balanced_partition as B
    from anemoi_core_b200.distributed import khop_edges as K

    g = torch.Generator().manual_seed(1234)
    fails, checks = [], 0

    def check(ok: bool, what: str) -> None:
        nonlocal checks
        checks += 1
        if not ok and len(fails) < 10:
            fails.append(what)

    for total in range(0, 40):
        for parts in (1, 2, 3, 5, 8, 64):
            check(B.get_balanced_partition_sizes(total, parts) == RB.get_balanced_partition_sizes(total, parts), f"balanced({total}, {parts})")
    for case in range(n_cases):
        n_src = int(torch.randint(1, 81, (1,), generator=g))
        n_dst = int(torch.randint(1, 81, (1,), generator=g))
        n_edges = int(torch.randint(1, 401, (1,), generator=g))
        shape = case % 5
        src = torch.randint(0, n_src, (n_edges,), generator=g)
        dst = torch.randint(0, n_dst, (n_edges,), generator=g)
        if shape == 1:  # every edge into one destination
            dst = torch.full((n_edges,), int(torch.randint(0, n_dst, (1,), generator=g)))
        elif shape == 2:  # destinations only in the first third: whole parts without an edge
            dst = torch.randint(0, max(n_dst // 3, 1), (n_edges,), generator=g)
        elif shape == 3:  # duplicate edges
            src, dst = src.repeat(2)[:n_edges], dst.repeat(2)[:n_edges]
            half = n_edges // 2
            src[half : 2 * half], dst[half : 2 * half] = src[:half], dst[:half]
        ei = torch.stack([src, dst])
        tag = f"case {case} (n_src {n_src}, n_dst {n_dst}, E {n_edges}, shape {shape})"
        s_ref, p_ref = RK.sort_edge_index_by_dst(ei.clone())
        s, p = K.sort_edge_index_by_dst(ei.clone())
        check(torch.equal(s, s_ref) and torch.equal(p, p_ref), f"{tag}: sort")
        check(K.is_edge_index_dst_sorted(s) and bool(RK.is_edge_index_dst_sorted(s)) and K.is_edge_index_dst_sorted(ei) == bool(RK.is_edge_index_dst_sorted(ei)),
              f"{tag}: is_sorted")  # fmt: skip
        x_src = torch.arange(n_src, dtype=torch.float32).view(-1, 1)
        x_dst = torch.arange(n_dst, dtype=torch.float32).view(-1, 1)
        ea = torch.arange(n_edges, dtype=torch.float32).view(-1, 1)
        for parts in sorted({1, 2, 3, 7, n_dst, int(torch.randint(1, 12, (1,), generator=g))}):
            if parts > n_dst:
                continue
            gp_ref = RK.build_graph_partition(s, parts, (n_src, n_dst))
            gp = K.build_graph_partition(s, parts, (n_src, n_dst))
            check(list(gp.dst_splits) == list(gp_ref.dst_splits) and list(gp.edge_splits) == list(gp_ref.edge_splits), f"{tag}: splits, {parts} parts")
            for cid in range(parts):
                (xs_c, xd_c), ea_c, ei_c, _ = gp_ref.materialise(cid, (x_src, x_dst), ea, s)
                (d0, d1), (e0, e1), connected, local = gp.materialise(cid, s)
                check(torch.equal(torch.arange(d0, d1), xd_c.view(-1).long()) and torch.equal(torch.arange(e0, e1), ea_c.view(-1).long())
                      and torch.equal(connected, xs_c.view(-1).long()) and torch.equal(local, ei_c), f"{tag}: part {cid} of {parts}")  # fmt: skip
        # the helper on its own (mapper.py:248-297 uses it on the full bipartite graph)
        _, ei_ref, ids_ref = RK._drop_unconnected_src_nodes(x_src, s.clone())
        ids, ei_loc = K.drop_unconnected_src_nodes(n_src, s)
        check(torch.equal(ids, ids_ref) and torch.equal(ei_loc, ei_ref), f"{tag}: drop_unconnected_src_nodes")
    return {"cases": n_cases, "checks": checks, "failures": fails, "reference": root}


if __name__ == "__main__":
    print(json.dumps(main(int(sys.argv[1]) if len(sys.argv) > 1 else 200)))
