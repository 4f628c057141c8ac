"""Persistent forward pass (one cooperative launch, the MLPs are its stages; graphs with no more tiles than SMs) against
the launch-per-MLP path of the same library: the arithmetic per tile is the same code, so the results must be bit equal -
plain forward, repeated calls (the grid-barrier counter is reset on the stream), the launch count, and a captured CUDA
graph of the call (a cooperative launch inside stream capture)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import mgn_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _problem(nx, ny, mps, seed=0):
    rng = np.random.default_rng(seed)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, 2)
    ps = (orc.init_params(cfg, seed=seed + 1, dtype=np.float64)
          + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    nf = rng.normal(size=(N, 9)).astype(np.float32)
    ef = rng.normal(size=(E, 3)).astype(np.float32)
    return ps, nf, ef, s, r


def _model(pkg, mps, persist):
    old = os.environ.get("MGN_FWD_PERSIST")
    os.environ["MGN_FWD_PERSIST"] = "2" if persist else "0"      # read once, at mgn_model_create; 2 = also when captured
    try:
        return pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    finally:
        if old is None:
            del os.environ["MGN_FWD_PERSIST"]
        else:
            os.environ["MGN_FWD_PERSIST"] = old


@pytest.mark.parametrize("nx,ny,mps", [(5, 4, 1), (12, 9, 3), (65, 29, 15), (40, 29, 2)])
def test_persistent_forward_is_bit_equal(pkg, nx, ny, mps):
    ps, nf, ef, s, r = _problem(nx, ny, mps)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    m_on, m_off = _model(pkg, mps, True), _model(pkg, mps, False)
    lib = pkg.load()

    def counted(model):
        n, nt, ms = C.c_int64(0), C.c_int64(0), C.c_float(0)
        assert lib.mgn_profile_begin(-1) == 0
        out = model.forward(graph, dev(ps), training=False)
        assert lib.mgn_profile_end(C.byref(n), C.byref(nt), C.byref(ms), None, 0) == 0
        return out, n.value

    a, n_on = counted(m_on)
    b, n_off = counted(m_off)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
    assert torch.equal(a, m_on.forward(graph, dev(ps), training=False))   # the barrier counter is reset per call
    assert n_off == 3 + 2 * mps + 1                                       # pack + one launch per MLP
    assert n_on == 2                                                      # pack + the persistent launch


def test_persistent_forward_in_a_cuda_graph(pkg):
    ps, nf, ef, s, r = _problem(65, 29, 15, seed=3)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    model = _model(pkg, 15, True)
    p = dev(ps)
    eager = model.forward(graph, p, training=False).clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model.forward(graph, p, training=False)       # warm-up on the capture stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            out = model.forward(graph, p, training=False)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
