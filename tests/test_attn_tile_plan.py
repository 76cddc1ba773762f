"""Host-side planner of the destination-tile attention kernel (anemoi_b200_attn_tile_plan, include/anemoi_b200.h): runs on the CPU, so
its invariants are checked here without a GPU.  Every edge must land in a slot of its tile that holds its source row, no two edges of one
destination may share a slot (duplicate (src, dst) pairs included), tiles respect the 16-row / 64-slot / 512-edge limits and tile the
destination range in order."""
import numpy as np
import pytest
import torch

from anemoi_core_b200 import ops
from anemoi_core_b200.synthetic import build_graph
from anemoi_core_b200.synthetic import random_graph


def colptr_of(ei: np.ndarray, n_dst: int) -> np.ndarray:
    c = np.zeros(n_dst + 1, np.int64)
    np.add.at(c, ei[1] + 1, 1)
    return np.cumsum(c).astype(np.int32)


def check_plan(src, colptr, n_src, n_dst):
    res = ops.plan_attention_tiles_host(src, colptr, n_src, n_dst)
    assert res is not None
    tm, ss, es = res
    E = len(src)
    nd, U, ne = tm[:, 3] & 0xFF, (tm[:, 3] >> 8) & 0xFF, tm[:, 3] >> 16
    assert tm[0, 0] == 0 and nd.sum() == n_dst and (nd >= 1).all() and (nd <= 16).all() and (U <= 64).all() and (ne <= ops.ATTN_TILE_MAX_EDGES).all()
    assert (np.cumsum(nd)[:-1] == tm[1:, 0]).all() and (np.cumsum(U)[:-1] == tm[1:, 1]).all() and U.sum() == len(ss)
    assert (tm[:, 2] == colptr[tm[:, 0]]).all() and (ne == colptr[tm[:, 0] + nd] - colptr[tm[:, 0]]).all()
    tile_of_dst = np.repeat(np.arange(len(tm)), nd)
    dst = np.repeat(np.arange(n_dst), np.diff(colptr))
    if E:
        assert ((es >> 8) == dst - tm[tile_of_dst[dst], 0]).all()  # row of the edge inside its tile
        es = es & 0xFF
        assert (es < U[tile_of_dst[dst]]).all()
        assert (ss[tm[tile_of_dst[dst], 1] + es] == src).all()  # the slot holds the edge's source row
        assert len(np.unique(dst.astype(np.int64) * 64 + es)) == E  # one edge per (destination, slot)
    return tm, ss, es


def test_plan_on_the_icosahedral_mesh_natural_and_locality_order():
    from anemoi_core_b200.layers import _reorder as RO

    gr = build_graph("o32", 4)
    ei, n = gr["proc_index"].numpy(), gr["n_mesh"]
    _, ss0, _ = check_plan(ei[0].astype(np.int32), colptr_of(ei, n), n, n)
    plan = RO.locality_plan(gr["proc_index"], n, min_nodes=0)
    e2 = plan.edge_index.numpy()
    tm, ss1, _ = check_plan(e2[0].astype(np.int32), colptr_of(e2, n), n, n)
    assert len(ss1) < 0.6 * len(ss0)  # the locality order shares gathered rows inside a tile
    assert (tm[:, 3] & 0xFF).mean() > 15


@pytest.mark.parametrize("seed", [1, 2])
def test_plan_on_random_and_bipartite_graphs(seed):
    rg = random_graph(500, 321, 4000, 3, seed=seed)["index"].numpy()
    check_plan(rg[0].astype(np.int32), colptr_of(rg, 321), 500, 321)


def test_plan_duplicates_empty_rows_and_limits():
    src = np.array([3, 3, 3, 5, 1, 1, 7], np.int32)
    cp = np.array([0, 4, 4, 6, 7, 7], np.int32)
    tm, ss, es = check_plan(src, cp, 8, 5)
    assert len(tm) == 1 and len(ss) == 7  # three slots for the triple edge 3 -> 0
    # no edges at all
    tm, ss, es = check_plan(np.zeros(0, np.int32), np.zeros(4, np.int32), 5, 3)
    assert len(tm) == 1 and len(ss) == 0
    # 64 distinct sources into one row still fit; 65 do not
    ok = ops.plan_attention_tiles_host(np.arange(64, dtype=np.int32), np.array([0, 64], np.int32), 64, 1)
    assert ok is not None and ok[0][0, 3] == (1 | (64 << 8) | (64 << 16))
    assert ops.plan_attention_tiles_host(np.arange(65, dtype=np.int32), np.array([0, 65], np.int32), 65, 1) is None
    # a row that does not fit the open tile starts the next one, and the rolled-back slots are not left behind
    src = np.concatenate([np.arange(40), np.arange(30, 70)]).astype(np.int32)
    tm, ss, es = check_plan(src, np.array([0, 40, 80], np.int32), 70, 2)
    assert len(tm) == 2 and list(tm[:, 3] & 0xFF) == [1, 1] and len(ss) == 80
    with pytest.raises(RuntimeError, match="out of range"):
        ops.plan_attention_tiles_host(np.array([9], np.int32), np.array([0, 1], np.int32), 3, 1)
