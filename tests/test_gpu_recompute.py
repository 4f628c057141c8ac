"""MGN_RECOMPUTE=1: the processor MLPs keep no activation saves; the backward pass re-runs each MLP (saves only) right before
its own kernels.  The recomputed saves are the same bits, so loss, parameter gradient and d/d(node features) must be bit
equal to the default path - and the training workspace must shrink."""
import os

import numpy as np
import pytest
import torch

import mgn_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(pkg, mps, recompute, **kw):
    old = os.environ.get("MGN_RECOMPUTE")
    os.environ["MGN_RECOMPUTE"] = "1" if recompute else "0"      # read once, at mgn_model_create
    try:
        return pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=pkg.COMPUTE_BF16, **kw)
    finally:
        if old is None:
            del os.environ["MGN_RECOMPUTE"]
        else:
            os.environ["MGN_RECOMPUTE"] = old


@pytest.mark.parametrize("nx,ny,mps,windows,post", [(12, 9, 3, 1, False), (65, 29, 15, 1, False), (30, 17, 4, 3, False),
                                                    (20, 11, 3, 1, True)])
def test_recompute_is_bit_equal_and_smaller(pkg, nx, ny, mps, windows, post):
    rng = np.random.default_rng(11)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    N0 = pos.shape[0]
    cells_b = np.concatenate([cells + b * N0 for b in range(windows)], axis=0)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells_b))
    N, E = N0 * windows, s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, 2)
    ps = (orc.init_params(cfg, seed=5, dtype=np.float64) + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    nf, ef = rng.normal(size=(N, 9)).astype(np.float32), rng.normal(size=(E, 3)).astype(np.float32)
    tgt, mask = rng.normal(size=(N, 2)).astype(np.float32), orc.node_mask(np.tile(nt, windows), [0, 5])
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    res = []
    for rec in (False, True):
        model = _model(pkg, mps, rec, aggregate_post_residual=post)
        mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
        (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
        out = model.forward(graph, dev(ps), training=True)
        dps, dnf = model.backward(graph, dev(ps), torch.ones_like(out), want_dnf=True)
        torch.cuda.synchronize()
        res.append((gs.clone(), loss.clone(), dps.clone(), dnf.clone(), model.workspace(graph.index, True).numel()))
    (g0, l0, p0, n0, w0), (g1, l1, p1, n1, w1) = res
    assert torch.isfinite(g0).all() and float(g0.abs().sum()) > 0
    assert torch.equal(l0, l1) and torch.equal(g0, g1) and torch.equal(p0, p1) and torch.equal(n0, n1)
    assert w1 < w0
    if mps >= 15:
        assert w1 < 0.55 * w0      # 15 MP steps: the saves of 30 MLPs collapse into two sets (one window: the backward
                                   # scratch - per-CTA weight-gradient partials - is the other half)


def test_recompute_workspace_of_the_bench_graph(pkg):
    """32 CylinderFlow windows, 15 MP steps (sizes only: no launch): the training workspace falls to about a third."""
    pos, cells, nt = orc.cylinder_flow_mesh(65, 29)
    N0, B = pos.shape[0], 32
    cells_b = np.concatenate([cells + b * N0 for b in range(B)], axis=0)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells_b))
    graph = pkg.FeatureGraph(torch.zeros((N0 * B, 9), device="cuda"), torch.zeros((s.shape[0], 3), device="cuda"), dev(s), dev(r))
    w0 = _model(pkg, 15, False).workspace_bytes(graph.index, True)
    w1 = _model(pkg, 15, True).workspace_bytes(graph.index, True)
    print(f"training workspace, 32 windows x 15 MP steps: {w0 / 1e9:.1f} GB -> {w1 / 1e9:.1f} GB with MGN_RECOMPUTE=1")
    assert w1 < 0.4 * w0
