"""Known-answer tests that pin the integer path.  The reference ships no golden
vectors (test/runtests.jl:10-18 is Aqua only), so these are hand-derived from the stated rules and
checked on BOTH the oracle and the library's host entry points (no GPU needed)."""
import numpy as np
import pytest

import mgn_oracle as orc


def test_one_hot_kat(pkg):
    want = np.zeros((3, 7), np.float32)
    want[0, 0] = want[1, 5] = want[2, 6] = 1  # rows v+offset = 1, 6, 7 (1-based)
    assert np.array_equal(orc.one_hot([0, 5, 6], 7, 1), want)
    assert np.array_equal(pkg.one_hot([0, 5, 6], 7, 1), want)


def test_triangles_to_edges_kat(pkg):
    # faces (0,1,2),(1,2,3): raw (0,1),(1,2)|(1,2),(2,3)|(2,0),(3,1) -> (max,min), unique first-occurrence
    cells = np.array([[0, 1, 2], [1, 2, 3]], np.int32)
    want_s = [1, 2, 3, 2, 3, 0, 1, 2, 0, 1]
    want_r = [0, 1, 2, 0, 1, 1, 2, 3, 2, 3]
    for impl in (orc.triangles_to_edges, pkg.triangles_to_edges):
        s, r = impl(cells)
        assert s.dtype == np.int32 and s.tolist() == want_s and r.tolist() == want_r
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))  # graph.jl:31-34
    assert s.tolist() == [2, 3, 4, 3, 4, 1, 2, 3, 1, 2] and r.tolist() == [1, 2, 3, 1, 2, 2, 3, 4, 3, 4]
    s2, r2 = pkg.triangles_to_edges(cells)
    assert pkg.shift_one_based(s2, r2) is True
    assert np.array_equal(s2, s) and np.array_equal(r2, r)
    assert pkg.shift_one_based(s2, r2) is False  # already 1-based: untouched
    rp, perm = orc.build_csr(r, 4)
    assert rp.tolist() == [0, 2, 5, 8, 10]
    assert (perm + 1).tolist() == [1, 4, 2, 5, 6, 3, 7, 9, 8, 10]


def test_parse_edges_chain_kat(pkg):
    e = orc.create_edges_1d(5)  # dataset.jl:379-382
    assert e.tolist() == [[1, 2], [2, 3], [3, 4], [4, 5]]
    for impl in (orc.parse_edges, pkg.parse_edges):
        s, r = impl(e)
        assert s.tolist() == [1, 2, 3, 4, 2, 3, 4, 5] and r.tolist() == [2, 3, 4, 5, 1, 2, 3, 4]
    rp, perm = orc.build_csr(r, 5)
    assert rp.tolist() == [0, 1, 3, 5, 7, 8]
    assert (perm + 1).tolist() == [5, 1, 6, 2, 7, 3, 8, 4]


def test_mask_and_mirror_property():
    pos, cells, nt = orc.cylinder_flow_mesh()
    assert pos.shape == (1885, 2) and cells.shape == (3584, 3)
    s, r = orc.triangles_to_edges(cells)
    E = s.shape[0]
    assert E == 10936
    assert np.array_equal(s, np.roll(r, E // 2))  # senders[k] == receivers[(k + E/2) mod E]
    m = orc.node_mask([0, 5, 6, 0, 4], [0, 5])
    assert m.tolist() == [1, 2, 4] and m.dtype == np.int32
    vm = orc.val_mask([0, 5, 6], [0, 5], 2)
    assert vm.tolist() == [[1, 1], [1, 1], [0, 0]]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_library_host_indexing_matches_oracle(pkg, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 60))
    cells = rng.integers(0, n, size=(int(rng.integers(1, 120)), 3)).astype(np.int32)
    so, ro = orc.triangles_to_edges(cells)
    sl, rl = pkg.triangles_to_edges(cells)
    assert np.array_equal(so, sl) and np.array_equal(ro, rl)
    so, ro = orc.shift_to_one_based(so, ro)
    pkg.shift_one_based(sl, rl)
    assert np.array_equal(so, sl) and np.array_equal(ro, rl)
    pos = rng.normal(size=(n, 2)).astype(np.float32)
    assert np.array_equal(orc.edge_features(pos, so, ro), pkg.edge_features(pos, sl, rl))  # bit-exact
    v = rng.integers(0, 7, size=40)
    assert np.array_equal(orc.one_hot(v, 7, 1), pkg.one_hot(v, 7, 1))


def test_empty_and_degenerate_inputs(pkg):
    s, r = pkg.triangles_to_edges(np.zeros((0, 3), np.int32))
    assert s.shape == (0,) and r.shape == (0,)
    s, r = pkg.parse_edges(np.zeros((0, 2), np.int32))
    assert s.shape == (0,)
    # a degenerate face (a, a, b) produces the self edge (a, a) once
    s, r = pkg.triangles_to_edges(np.array([[2, 2, 5]], np.int32))
    so, ro = orc.triangles_to_edges(np.array([[2, 2, 5]], np.int32))
    assert np.array_equal(s, so) and np.array_equal(r, ro) and s.tolist() == [2, 5, 2, 2]
    with pytest.raises(pkg.MgnError):
        pkg.edge_features(np.zeros((2, 2), np.float32), np.array([3], np.int32), np.array([1], np.int32))


def test_normalisers_oracle():
    n = orc.NormaliserOnline(2)
    x1 = np.array([[1.0, 10.0], [3.0, 30.0]], np.float32)
    y = n(x1)
    # mean = (2, 20), var = (1, 100)
    assert np.allclose(y, [[-1, -1], [1, 1]])
    assert n.acc_count == 2 and n.num_acc == 1
    assert np.allclose(n.inverse(y), x1)
    n2 = orc.NormaliserOnline(1, max_acc=1)
    n2(np.array([[2.0], [4.0]], np.float32))
    n2(np.array([[100.0]], np.float32))  # num_acc reached max_acc: no more accumulation
    assert n2.acc_count == 2
    mm = orc.NormaliserOfflineMinMax(0.0, 6.0)
    assert np.allclose(mm(np.array([[3.0]])), 0.5) and np.allclose(mm.inverse(np.array([[0.5]])), 3.0)
    ms = orc.NormaliserOfflineMeanStd(1.0, 2.0)
    assert np.allclose(ms(np.array([[5.0]])), 2.0) and np.allclose(ms.inverse(np.array([[2.0]])), 5.0)


def test_product_workload_generators_match_the_oracle(pkg):
    """bench.py / tools build their inputs with the product-side generators (meshgraphnets.jl_b200/workloads.py);
    they must be the same workloads the oracle defines."""
    import mgn_oracle as orc
    for a, b in zip(pkg.cylinder_flow_mesh(65, 29), orc.cylinder_flow_mesh(65, 29)):
        assert np.array_equal(a, b)
    pos, _, nt = orc.cylinder_flow_mesh(9, 7)
    assert np.array_equal(pkg.synthetic_velocity(pos, 5, seed=3), orc.synthetic_velocity(pos, 5, seed=3))
    assert np.array_equal(pkg.node_mask(nt, [0, 5]), orc.node_mask(nt, [0, 5]))
    assert np.array_equal(pkg.val_mask(nt, [0, 5], 2), orc.val_mask(nt, [0, 5], 2))
    assert np.array_equal(pkg.chain_edges(17), orc.create_edges_1d(17))
    e = pkg.tet_grid_edges(5)
    assert e.shape == (3 * 5 * 5 * 4 + 3 * 5 * 4 * 4 + 4 ** 3, 2) and (e[:, 0] < e[:, 1]).all()
    assert np.array_equal(e, np.unique(e, axis=0))          # sorted, unique
