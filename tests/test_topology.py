"""Unit tests of the host map topology (no GPU, no oracle)."""
import networkx as nx
import numpy as np
import pytest

from dbgsom_b200.topology import HOP_INF, MapTopology, RowOp


def grown_map(seed=0, steps=12):
    rng = np.random.default_rng(seed)
    t = MapTopology.initial_square()
    for epoch in range(steps):
        t.error[:] = rng.uniform(0, 10, len(t))
        t.distribute_errors(6.0)
        t.grow(6.0, epoch)
    return t


def test_initial_square_matches_reference_layout():
    t = MapTopology.initial_square()
    assert t.pos == [(0, 0), (0, 1), (1, 0), (1, 1)]
    g = t.to_networkx()
    assert list(g.nodes) == t.pos
    assert sorted(map(sorted, g.edges)) == [[(0, 0), (0, 1)], [(0, 0), (1, 0)], [(0, 1), (1, 1)], [(1, 0), (1, 1)]]
    assert [list(g.adj[p]) for p in t.pos] == [[t.pos[j] for j in a] for a in t.adj]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_hop_matrix_equals_floyd_warshall(seed):
    t = grown_map(seed)
    assert len(t) > 8
    g = t.to_networkx()
    np.testing.assert_array_equal(t.hop_matrix(), nx.floyd_warshall_numpy(g))
    u = t.hop_matrix_u16()
    assert u.dtype == np.uint16 and (u != HOP_INF).all()


def test_export_keeps_adjacency_order():
    t = grown_map(5, 15)
    g = t.to_networkx()
    for i, p in enumerate(t.pos):
        assert list(g.adj[p]) == [t.pos[j] for j in t.adj[i]]


def test_hops_after_removal_can_be_infinite():
    t = MapTopology.full_grid(1, 5)
    cut = t.without(np.array([2]))
    d = cut.hop_matrix()
    assert np.isinf(d[0, 3]) and d[0, 1] == 1 and d[2, 3] == 1
    assert cut.hop_matrix_u16()[0, 3] == HOP_INF
    g = t.to_networkx()
    g.remove_node((0, 2))
    np.testing.assert_array_equal(d, nx.floyd_warshall_numpy(g))


def test_full_grid_matches_networkx_grid():
    t = MapTopology.full_grid(4, 3)
    g = nx.grid_2d_graph(4, 3)
    assert t.pos == list(g.nodes)
    for i, p in enumerate(t.pos):
        assert list(g.adj[p]) == [t.pos[j] for j in t.adj[i]]
    np.testing.assert_array_equal(t.hop_matrix(), nx.floyd_warshall_numpy(g))


def test_row_ops_reference_extrapolation_and_overwrite():
    t = MapTopology.initial_square()
    t.error[:] = [10.0, 1.0, 2.0, 3.0]
    ops = t.grow(5.0, epoch=7)
    # node (0,0) has degree 2 with neighbours (0,1) [err 1] and (1,0) [err 2]: not e1 > e2, so the
    # new node mirrors nbr1=(0,1) through (0,0) -> (0,-1), W = 2 W[(0,0)] - W[(0,1)]
    assert ops == [RowOp(dst=4, a=0, b=1, c=-1)]
    assert t.pos[4] == (0, -1) and t.epoch_created[4] == 7 and t.adj[4] == [0]
    # placing on an occupied cell keeps the row, resets error, updates the epoch
    t.error[4] = 3.0
    assert t.place((0, -1), epoch=9) == 4
    assert t.error[4] == 0.0 and t.epoch_created[4] == 9 and len(t) == 5


def test_shadow_row_permutation_is_a_bijection():
    """engine.scatter_stride: (c * stride) % mpad must visit every shadow row once (the candidate kernel evaluates
    this formula in registers, dbgsom_bmu_args.proto_stride) and scatter neighbouring rows across the map."""
    import numpy as np

    from dbgsom_b200.engine import scatter_stride

    for mpad in (256, 512, 768, 1024, 4096, 5120, 16384, 24576, 65280):
        s = scatter_stride(mpad)
        assert 0 < s < mpad
        proto = (np.arange(mpad, dtype=np.int64) * s) % mpad
        assert np.array_equal(np.sort(proto), np.arange(mpad))
        assert int(mpad) * s < 2**32  # the kernel multiplies in 32 bits
        # consecutive shadow rows are far apart on the map
        assert np.abs(np.diff(proto)).min() > mpad // 4
