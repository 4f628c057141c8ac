"""Property tests (hypothesis) of the integer path and of the oracle (SURVEY.md section 4, item iv).

The host half of the C ABI (one_hot, triangles_to_edges, parse_edges, the 0 -> 1 shift, edge features) runs without a
GPU, so the LIBRARY is compared bit for bit with the oracle on generated inputs - ragged, with duplicate faces,
degenerate triangles, isolated nodes and repeated edges.  The oracle itself is checked for the structural properties
the kernels rely on: the CSR is a stable sort (a permutation, receivers non-decreasing, edge ids ascending inside a
segment), the sequential scatter equals the segmented sum over it, the model output does not depend on the edge order,
and the pullback is linear in its cotangent."""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import mgn_oracle as orc

SET = settings(max_examples=40, deadline=None)


@st.composite
def cell_arrays(draw):
    n_nodes = draw(st.integers(3, 40))
    n_cells = draw(st.integers(0, 60))
    base = draw(st.sampled_from([0, 1]))
    cells = draw(st.lists(st.tuples(*[st.integers(base, n_nodes - 1 + base)] * 3), min_size=n_cells, max_size=n_cells))
    return np.asarray(cells, dtype=np.int32).reshape(-1, 3), n_nodes, base


@SET
@given(cell_arrays())
def test_triangles_to_edges_library_equals_oracle(pkg, ca):
    cells, n_nodes, base = ca
    s, r = pkg.triangles_to_edges(cells)
    s_o, r_o = orc.triangles_to_edges(cells)
    assert np.array_equal(s, s_o) and np.array_equal(r, r_o) and s.dtype == np.int32
    U = s.shape[0] // 2
    # two-way: the second half is the first half mirrored; (max, min) convention; unique undirected edges
    assert np.array_equal(s[:U], r[U:]) and np.array_equal(r[:U], s[U:])
    assert (s[:U] >= r[:U]).all()
    assert len({(int(a), int(b)) for a, b in zip(s[:U], r[:U])}) == U
    # the in-place shift fires exactly when a zero id is present (src/graph.jl:31-34)
    s2, r2 = s.copy(), r.copy()
    shifted = pkg.shift_one_based(s2, r2)
    s_o2, r_o2 = orc.shift_to_one_based(s_o, r_o)
    assert shifted == bool(((s == 0) | (r == 0)).any())
    assert np.array_equal(s2, s_o2) and np.array_equal(r2, r_o2)


@SET
@given(st.lists(st.tuples(st.integers(1, 50), st.integers(1, 50)), min_size=0, max_size=80))
def test_parse_edges_library_equals_oracle(pkg, pairs):
    e = np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    s, r = pkg.parse_edges(e)
    s_o, r_o = orc.parse_edges(e)
    assert np.array_equal(s, s_o) and np.array_equal(r, r_o)
    assert np.array_equal(s[:len(pairs)], e[:, 0]) and np.array_equal(s[len(pairs):], e[:, 1])


@SET
@given(st.lists(st.integers(-3, 12), min_size=0, max_size=50), st.integers(1, 9), st.integers(-2, 3))
def test_one_hot_library_equals_oracle(pkg, v, depth, offset):
    got = pkg.one_hot(np.asarray(v, np.int32), depth, offset)
    want = orc.one_hot(np.asarray(v, np.int32), depth, offset)
    assert np.array_equal(got, want) and got.dtype == np.float32
    assert ((got.sum(axis=1) == 1) | (got.sum(axis=1) == 0)).all()       # out-of-range rows stay zero


@SET
@given(st.integers(1, 3), st.integers(2, 30), st.integers(0, 60), st.integers(0, 2 ** 31 - 1))
def test_edge_features_library_equals_oracle_bitwise(pkg, dim, n_nodes, n_edges, seed):
    rng = np.random.default_rng(seed)
    pos = (rng.normal(size=(n_nodes, dim)) * 10.0 ** rng.integers(-3, 4)).astype(np.float32)
    s = rng.integers(1, n_nodes + 1, size=n_edges).astype(np.int32)
    r = rng.integers(1, n_nodes + 1, size=n_edges).astype(np.int32)
    got = pkg.edge_features(pos, s, r)
    want = orc.edge_features(pos, s, r)
    assert got.shape == (n_edges, dim + 1) and np.array_equal(got, want)      # fp32 subtraction, widened norm


@SET
@given(st.integers(1, 30), st.integers(0, 200), st.integers(0, 2 ** 31 - 1))
def test_csr_is_a_stable_sort_and_scatter_equals_segmented_sum(n_nodes, n_edges, seed):
    rng = np.random.default_rng(seed)
    recv = rng.integers(1, n_nodes + 1, size=n_edges).astype(np.int32)
    rp, perm = orc.build_csr(recv, n_nodes)
    assert rp[0] == 0 and rp[-1] == n_edges and (np.diff(rp) >= 0).all()
    assert sorted(perm.tolist()) == list(range(n_edges))
    for v in range(n_nodes):
        seg = perm[rp[v]:rp[v + 1]]
        assert (recv[seg] == v + 1).all() and (np.diff(seg) > 0).all()        # stable: ascending original edge id
    m = rng.normal(size=(n_edges, 3)).astype(np.float32)
    agg = orc.scatter_add(m, recv.astype(np.int64) - 1, n_nodes)
    seg = np.zeros_like(agg)
    for v in range(n_nodes):
        acc = np.zeros(3, np.float32)
        for j in range(rp[v], rp[v + 1]):
            acc = acc + m[perm[j]]
        seg[v] = acc
    assert np.array_equal(agg, seg)                                           # same summation order: same bits


def _tiny(seed):
    rng = np.random.default_rng(seed)
    pos, cells, nt = orc.cylinder_flow_mesh(4, 3)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    cfg = orc.ModelConfig(4, 3, 2, 8, 2, 1)
    p = orc.init_params(cfg, seed=seed % 1000, dtype=np.float64) + 0.05 * rng.normal(size=orc.mlp_specs(cfg)[1])
    N, E = pos.shape[0], s.shape[0]
    return cfg, p, rng.normal(size=(N, 4)), rng.normal(size=(E, 3)), s, r, rng


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_model_output_is_invariant_under_edge_relabelling(seed):
    cfg, p, nf, ef, s, r, rng = _tiny(seed)
    perm = rng.permutation(s.shape[0])
    a = orc.model_forward(cfg, p, nf, ef, s, r)
    b = orc.model_forward(cfg, p, nf, ef[perm], s[perm], r[perm])
    assert np.allclose(a, b, rtol=1e-11, atol=1e-13)


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.floats(-3, 3), st.floats(-3, 3))
def test_pullback_is_linear_in_the_cotangent(seed, a, b):
    cfg, p, nf, ef, s, r, rng = _tiny(seed)
    N = nf.shape[0]
    d1, d2 = rng.normal(size=(N, 2)), rng.normal(size=(N, 2))

    def vjp(d):
        tape = []
        orc.model_forward(cfg, p, nf, ef, s, r, 1, tape)
        return orc.model_backward(cfg, p, tape, d, s, r, N)
    g1, x1 = vjp(d1)
    g2, x2 = vjp(d2)
    g, x = vjp(a * d1 + b * d2)
    scale = max(np.abs(g1).max(), np.abs(g2).max(), 1e-30)
    assert np.allclose(g, a * g1 + b * g2, rtol=1e-9, atol=1e-11 * scale)
    assert np.allclose(x, a * x1 + b * x2, rtol=1e-9, atol=1e-11 * max(np.abs(x1).max(), np.abs(x2).max(), 1e-30))


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 40), st.integers(0, 150), st.integers(1, 9), st.integers(0, 2 ** 31 - 1))
def test_partition_plan_invariants_on_random_multigraphs(pkg, n_nodes, n_edges, world, seed):
    """The halo-exchange plan (meshgraphnets.jl_b200/partition.py, integer work): random multigraphs with self loops,
    repeated edges, isolated nodes and more ranks than nodes.  Every edge has exactly one owner (the owner of its
    receiver), local ids map back to the global graph, halo rows are exactly the remote senders, the send / receive
    lists of every pair of ranks describe the same rows in the same order, and the one-rank builder agrees."""
    rng = np.random.default_rng(seed)
    s = rng.integers(1, n_nodes + 1, size=n_edges).astype(np.int32)
    r = rng.integers(1, n_nodes + 1, size=n_edges).astype(np.int32)
    parts = pkg.build_partition(n_nodes, s, r, world)
    assert [p.lo for p in parts] + [parts[-1].hi] == pkg.partition_bounds(n_nodes, world)
    all_e = np.concatenate([p.edge_ids for p in parts]) if parts else np.zeros(0, np.int64)
    assert np.array_equal(np.sort(all_e), np.arange(n_edges))
    for p in parts:
        glob = p.local_nodes_global()
        assert np.array_equal(glob[p.senders - 1], s[p.edge_ids] - 1)
        assert np.array_equal(glob[p.receivers - 1], r[p.edge_ids] - 1)
        assert ((p.receivers - 1) < p.n_own).all() and (np.diff(p.edge_ids) > 0).all()
        remote = np.unique(s[p.edge_ids][(s[p.edge_ids] - 1 < p.lo) | (s[p.edge_ids] - 1 >= p.hi)] - 1)
        assert np.array_equal(p.halo_global, remote)
        for q, rows in p.recv_rows.items():
            assert np.array_equal(parts[q].send_rows[p.rank] + parts[q].lo, p.halo_global[rows - p.n_own])
        covered = np.sort(np.concatenate(list(p.recv_rows.values()))) if p.recv_rows else np.zeros(0, np.int64)
        assert np.array_equal(covered, np.arange(p.n_own, p.n_local))
        one = pkg.build_partition_rank(n_nodes, s, r, world, p.rank)
        assert np.array_equal(one.senders, p.senders) and np.array_equal(one.receivers, p.receivers)
        assert np.array_equal(one.halo_global, p.halo_global) and np.array_equal(one.edge_ids, p.edge_ids)
        assert set(one.send_rows) == set(p.send_rows) and set(one.recv_rows) == set(p.recv_rows)
        assert all(np.array_equal(one.send_rows[q], p.send_rows[q]) for q in p.send_rows)
        assert all(np.array_equal(one.recv_rows[q], p.recv_rows[q]) for q in p.recv_rows)
