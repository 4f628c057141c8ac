"""create_base_graph on the device (src/graph.jl:25-55): the hash-set unique of triangles_to_edges,
parse_edges, the 0 -> 1 shift, one_hot and the edge features, bit-exact against the oracle and the host C path."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _both(pkg, data_h, type_size=6, type_min=0):
    host = pkg.create_base_graph(data_h, type_size, type_min)
    data_d = {k: dev(np.asarray(v)) for k, v in data_h.items()}
    devr = pkg.create_base_graph(data_d, type_size, type_min)          # CUDA tensors -> device path
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in host], [t.cpu().numpy() for t in devr]


def test_survey_worked_example(pkg):
    """faces (0,1,2),(1,2,3) -> senders [2,3,4,3,4,1,2,3,1,2], receivers [1,2,3,1,2,2,3,4,3,4] after the shift."""
    data = {"node_type": np.array([0, 4, 5, 6], np.int32).reshape(1, -1, 1),
            "mesh_pos": np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)[None],
            "cells": np.array([[0, 1, 2], [1, 2, 3]], np.int32)[None]}
    host, d = _both(pkg, data)
    assert d[1].tolist() == [2, 3, 4, 3, 4, 1, 2, 3, 1, 2] and d[2].tolist() == [1, 2, 3, 1, 2, 2, 3, 4, 3, 4]
    for a, b in zip(host, d):
        assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("nx,ny,batch", [(65, 29, 1), (13, 9, 5), (65, 29, 32)])
def test_cylinder_flow_meshes_bit_exact(pkg, nx, ny, batch):
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    N = pos.shape[0]
    data = {"node_type": np.tile(nt, batch).reshape(1, -1, 1), "mesh_pos": np.tile(pos, (batch, 1))[None],
            "cells": np.concatenate([cells + b * N for b in range(batch)], axis=0)[None]}
    host, d = _both(pkg, data)
    s_o, r_o = orc.shift_to_one_based(*orc.triangles_to_edges(data["cells"][0]))
    assert np.array_equal(d[1], s_o) and np.array_equal(d[2], r_o)
    assert np.array_equal(d[0], orc.one_hot(data["node_type"].reshape(-1), 7, 1))
    assert np.array_equal(d[3], orc.edge_features(data["mesh_pos"][0], s_o, r_o))     # fp32 rel, widened norm: same bits
    for a, b in zip(host, d):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("seed", range(4))
def test_random_triangle_soup_first_occurrence_order(pkg, seed):
    """Duplicated edges within and across faces, degenerate faces (a == b), already 1-based ids (no shift): the unique
    must keep the FIRST occurrence in the [f0f1 ; f1f2 ; f2f0] order."""
    rng = np.random.default_rng(seed)
    n, C = 40, 300
    base = seed % 2                                # odd seeds: ids start at 1 -> no shift
    cells = rng.integers(base, n + base, size=(C, 3)).astype(np.int32)
    data = {"node_type": rng.integers(0, 7, size=n + base).astype(np.int32).reshape(1, -1, 1),
            "mesh_pos": rng.normal(size=(1, n + base + 1, 3)).astype(np.float32), "cells": cells[None]}
    host, d = _both(pkg, data)
    s_o, r_o = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    assert np.array_equal(d[1], s_o) and np.array_equal(d[2], r_o)
    for a, b in zip(host, d):
        assert np.array_equal(a, b)


def test_edge_list_chain_bit_exact(pkg):
    """BASELINE configs[3] entry: src/dataset.jl:379-382 chain edges through parse_edges (1-D positions, F_e = 2)."""
    n = 5000
    edges = orc.create_edges_1d(n)
    data = {"node_type": np.zeros((1, n, 1), np.int32), "mesh_pos": np.linspace(0, 1, n, dtype=np.float32).reshape(1, n, 1),
            "edges": edges}
    host, d = _both(pkg, data)
    s_o, r_o = orc.parse_edges(edges)
    assert np.array_equal(d[1], s_o) and np.array_equal(d[2], r_o)
    for a, b in zip(host, d):
        assert np.array_equal(a, b)


def test_device_path_rejects_out_of_range_ids(pkg):
    pos = dev(np.zeros((3, 2), np.float32))
    s, r = dev(np.array([1, 9], np.int32)), dev(np.array([2, 1], np.int32))
    out = torch.empty((2, 3), device="cuda")
    import ctypes as C
    with pytest.raises(pkg.MgnError) as e:
        pkg._lib.call("mgn_edge_features_device", C.c_void_p(pos.data_ptr()), 3, 2, C.c_void_p(s.data_ptr()),
                      C.c_void_p(r.data_ptr()), 2, 1, C.c_void_p(out.data_ptr()), None)
    assert e.value.code == 3
