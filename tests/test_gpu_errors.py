"""Error behaviour of the C ABI on the device path: status codes + mgn_last_error, never an abort
(include/mgn_b200.h conventions; the Julia shim turns a non-zero status into error(...))."""
import ctypes as C

import numpy as np
import pytest
import torch

import mgn_oracle as orc

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _graph(pkg, nx=6, ny=5):
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    rng = np.random.default_rng(0)
    g = pkg.FeatureGraph(dev(rng.normal(size=(pos.shape[0], 9)).astype(np.float32)),
                         dev(rng.normal(size=(s.shape[0], 3)).astype(np.float32)), dev(s), dev(r))
    return g


def test_workspace_too_small_is_reported(pkg):
    from meshgraphnets_jl_b200.core import _ptr, _stream
    g = _graph(pkg)
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        model, ps, _ = pkg.build_model(9, 2, 2, 2, 128, 2, compute_mode=mode)
        out = torch.empty(g.node_features.shape[0], 2, device="cuda")
        small = torch.empty(1024, dtype=torch.uint8, device="cuda")
        st = pkg.load().mgn_forward(model._h, g.index._h, _ptr(ps), _ptr(g.node_features), _ptr(g.edge_features),
                                    _ptr(out), _ptr(small), small.numel(), 0, _stream())
        from meshgraphnets_jl_b200._lib import last_error
        assert st == 4 and "workspace" in last_error()


def test_bad_stage_and_bad_halo_arguments(pkg):
    g = _graph(pkg)
    model, ps, _ = pkg.build_model(9, 2, 2, 2, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    with pytest.raises(pkg.MgnError) as e:
        model.forward_stage(g, ps, 7)                      # stage >= mps
    assert e.value.code == 1
    with pytest.raises(pkg.MgnError) as e:
        model.forward_stage(g, ps, -5)
    assert e.value.code == 1
    rows = torch.zeros(4, dtype=torch.int32, device="cuda")
    buf = torch.empty(4 * 256, dtype=torch.uint8, device="cuda")
    with pytest.raises(pkg.MgnError):                      # the bf16 latent cannot be accumulated
        model.halo_rows(g, True, pkg.HALO_LATENT, 0, rows, buf, pkg.ROWS_ADD)
    with pytest.raises(pkg.MgnError):                      # gradients need a training workspace
        model.halo_rows(g, False, pkg.HALO_GRAD, 0, rows, buf, pkg.ROWS_PACK)
    assert model.halo_row_bytes(pkg.HALO_LATENT) == 256 and model.halo_row_bytes(pkg.HALO_GRAD) == 512


def test_hub_node_is_unsupported_in_bf16_mode_but_fine_in_fp32(pkg):
    """A node with more than 128 incoming edges cannot be packed into a node-aligned tile: the tensor-core mode says
    so (MGN_ERR_UNSUPPORTED, code 5); the fp32 mode has no such limit."""
    n, e = 300, 400
    rng = np.random.default_rng(1)
    s = rng.integers(1, n + 1, size=e).astype(np.int32)
    r = rng.integers(1, n + 1, size=e).astype(np.int32)
    r[:200] = 5
    g = pkg.FeatureGraph(dev(rng.normal(size=(n, 9)).astype(np.float32)), dev(rng.normal(size=(e, 3)).astype(np.float32)),
                         dev(s), dev(r))
    m16, ps, _ = pkg.build_model(9, 2, 2, 1, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    with pytest.raises(pkg.MgnError) as err:
        m16.forward(g, ps)
    assert err.value.code == 5
    m32, ps32, _ = pkg.build_model(9, 2, 2, 1, 128, 2, compute_mode=pkg.COMPUTE_FP32)
    assert torch.isfinite(m32.forward(g, ps32)).all()


def test_bf16_mode_rejects_other_latent_sizes(pkg):
    with pytest.raises(pkg.MgnError) as err:
        pkg.Model(9, 3, 2, 2, 64, 2, compute_mode=pkg.COMPUTE_BF16)
    assert err.value.code == 1
