"""Tensor-core mode (MGN_COMPUTE_BF16: bf16 operands, fp32 accumulation in TMEM, fp32 master latents)
against the fp64 CPU oracle, through the C ABI.  Tolerances are the bf16-mode tolerances stated in
DESIGN.md: 3e-2 relative L2 on outputs after the residual stack, 6e-2 on gradients."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc
import mgn_oracle_bf16 as ob

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

TOL_OUT = 3e-2
TOL_GRAD = 6e-2


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _problem(nx, ny, mps, node_in=9, edge_in=3, seed=0, hidden=2):
    rng = np.random.default_rng(seed)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(node_in, edge_in, 2, 128, mps, hidden)
    ps = (orc.init_params(cfg, seed=seed + 1, dtype=np.float64)
          + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    nf = rng.normal(size=(N, node_in)).astype(np.float32)
    ef = rng.normal(size=(E, edge_in)).astype(np.float32)
    tgt = rng.normal(size=(N, 2)).astype(np.float32)
    mask = orc.node_mask(nt, [0, 5])
    return cfg, ps, nf, ef, s, r, tgt, mask


@pytest.mark.parametrize("nx,ny,mps,hidden", [(5, 4, 1, 2), (12, 9, 3, 2), (30, 17, 2, 1), (9, 6, 2, 0)])
def test_forward_bf16_matches_oracle(pkg, nx, ny, mps, hidden):
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(nx, ny, mps, hidden=hidden)
    out_o = orc.model_forward(cfg, ps.astype(np.float64), nf, ef, s, r, dtype=np.float64)
    model = pkg.Model(9, 3, 2, mps, 128, hidden, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    a = model.forward(graph, dev(ps), training=False)
    b = model.forward(graph, dev(ps), training=True)
    torch.cuda.synchronize()
    assert rel(a.cpu().numpy(), out_o) < TOL_OUT
    assert torch.equal(a, b)                       # inference and training forwards are the same arithmetic
    assert torch.equal(a, model.forward(graph, dev(ps), training=False))   # deterministic (no atomics)


def test_forward_bf16_full_size(pkg):
    """BASELINE configs[1] at full size: N=1885, E=10936, 15 MP steps."""
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(65, 29, 15)
    out_o = orc.model_forward(cfg, ps, nf, ef, s, r, dtype=np.float32)
    model = pkg.Model(9, 3, 2, 15, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    a = model.forward(graph, dev(ps), training=False)
    assert rel(a.cpu().numpy(), out_o) < TOL_OUT


def test_bf16_agrees_with_fp32_mode_on_device(pkg):
    """Property at a size the oracle does not reach in seconds: the two compute modes of the library
    agree within the bf16 tolerance on a 60k-edge mesh."""
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(120, 85, 4)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    m32 = pkg.Model(9, 3, 2, 4, 128, 2, compute_mode=pkg.COMPUTE_FP32)
    m16 = pkg.Model(9, 3, 2, 4, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    a = m32.forward(graph, dev(ps))
    b = m16.forward(graph, dev(ps))
    assert rel(b.cpu().numpy(), a.cpu().numpy()) < TOL_OUT


@pytest.mark.parametrize("nx,ny,mps,hidden", [(12, 9, 3, 2), (9, 7, 2, 1), (7, 5, 2, 0), (40, 30, 2, 2)])
def test_step_bf16_matches_oracle(pkg, nx, ny, mps, hidden):
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(nx, ny, mps, hidden=hidden)
    g_o, loss_o, out_o, dnf_o = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    model = pkg.Model(9, 3, 2, mps, 128, hidden, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
    try:
        (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
    except pkg.MgnError as e:
        if e.code == 5:
            pytest.skip("bf16 backward not built yet")
        raise
    assert abs(float(loss.cpu()) - loss_o) < TOL_OUT * abs(loss_o)
    assert rel(gs.cpu().numpy(), g_o) < TOL_GRAD
    for name, off, rows, cols in model.param_layout():
        ref = g_o[off:off + rows * cols]
        got = gs[off:off + rows * cols].cpu().numpy()
        assert np.linalg.norm(got - ref) <= 0.15 * np.linalg.norm(ref) + 1e-3 * np.linalg.norm(g_o), name
    out = model.forward(graph, dev(ps), training=True)
    _, dout_o = orc.loss_and_dout(out_o, tgt.astype(np.float64), mask)
    dps, dnf = model.backward(graph, dev(ps), dev(dout_o.astype(np.float32)), want_dnf=True)
    assert rel(dnf.cpu().numpy(), dnf_o) < 0.2   # raw-feature VJP: the end of the chain, all bf16 roundings accumulated


def test_step_bf16_many_tiles_per_cta(pkg):
    """More tiles than SMs (every CTA accumulates several tiles' weight gradients in TMEM): bf16 mode
    against the library's fp32 mode on a 60k-edge mesh; also run twice - bitwise deterministic."""
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(120, 85, 2)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    res = {}
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        model = pkg.Model(9, 3, 2, 2, 128, 2, compute_mode=mode)
        mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
        (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
        res[mode] = (gs.clone(), float(loss.cpu()))
        if mode == pkg.COMPUTE_BF16:
            (gs2,), _ = pkg.step_(mgn, graph, dev(tgt), dev(mask))
            assert torch.equal(gs, gs2)
    a, b = res[pkg.COMPUTE_FP32], res[pkg.COMPUTE_BF16]
    assert abs(a[1] - b[1]) < TOL_OUT * abs(a[1])
    assert rel(b[0].cpu().numpy(), a[0].cpu().numpy()) < TOL_GRAD


def test_bf16_kernels_reproduce_the_bf16_arithmetic_model(pkg):
    """The tight half of the bf16-mode parity claim: the tcgen05 kernels against the CPU model of the
    SAME arithmetic (oracle/mgn_oracle_bf16.py: the reference algorithm with a bf16 rounding wherever
    the kernels store bf16).  The only difference is the fp32 accumulation order, which flips the
    bf16 rounding of about one stored value in a few thousand; a flip perturbs its row by <= 2^-8 and
    then spreads one hop per MP step.  So: encoder + decoder (no message passing) agree to 1e-5; after
    one MP step >= 85 % of the nodes still agree to 1e-5 and none is off by more than 1e-2; the
    gradients of the 0-step model agree to 2e-3 and those of a 3-step model to 2e-2."""
    # (a) no message passing: two MLP chains, forward and backward
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(12, 9, 0)
    g_b, loss_b, out_b, dnf_b = ob.step_bf16(cfg, ps, nf, ef, s, r, tgt, mask)
    model = pkg.Model(9, 3, 2, 0, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    out = model.forward(graph, dev(ps), training=True)
    assert rel(out.cpu().numpy(), out_b) < 1e-5
    mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
    (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
    assert abs(float(loss.cpu()) - loss_b) < 1e-5 * abs(loss_b)
    assert rel(gs.cpu().numpy(), g_b) < 2e-3
    # (b) one MP step: most nodes bit-compatible, the rest within one bf16 ulp of their magnitude
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(12, 9, 1)
    out_b = ob.forward_bf16(cfg, ps, nf, ef, s, r)
    model = pkg.Model(9, 3, 2, 1, 128, 2, compute_mode=pkg.COMPUTE_BF16)
    out = model.forward(pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r)), dev(ps)).cpu().numpy()
    err = np.abs(out - out_b).max(axis=1) / np.abs(out_b).max()
    assert (err < 1e-5).mean() >= 0.85 and err.max() < 1e-2


@pytest.mark.parametrize("nx,ny,mps,hidden", [(12, 9, 3, 2), (9, 7, 2, 1), (7, 5, 2, 0)])
def test_step_bf16_close_to_bf16_arithmetic_model(pkg, nx, ny, mps, hidden):
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(nx, ny, mps, hidden=hidden)
    g_b, loss_b, out_b, dnf_b = ob.step_bf16(cfg, ps, nf, ef, s, r, tgt, mask)
    model = pkg.Model(9, 3, 2, mps, 128, hidden, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    out = model.forward(graph, dev(ps), training=True)
    assert rel(out.cpu().numpy(), out_b) < 1e-2
    mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
    (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
    assert abs(float(loss.cpu()) - loss_b) < 2e-3 * abs(loss_b)
    assert rel(gs.cpu().numpy(), g_b) < 3e-2      # 2.0e-2 observed: the bf16 edge-gradient stream adds rounding flips


def test_chain_100k_nodes_bf16_vs_fp32(pkg):
    """BASELINE configs[3] shape: 1-D chain of 100 000 nodes (src/dataset.jl:379-382 edges through parse_edges,
    E = 199 998, in-degree <= 2, F_e = 2): the CSR is bit-exact against the oracle and the two compute modes agree
    within the bf16 tolerance on outputs, loss and gradients (2 MP steps keep the test short)."""
    n = 100_000
    rng = np.random.default_rng(9)
    s, r = orc.parse_edges(orc.create_edges_1d(n))
    gi = pkg.GraphIndex(n, dev(s), dev(r))
    rp, perm, _, _ = gi.index_arrays()
    rp_o, perm_o = orc.build_csr(r, n)
    assert np.array_equal(rp, rp_o) and np.array_equal(perm, perm_o)
    cfg = orc.ModelConfig(3, 2, 1, 128, 2, 2)
    ps = dev((orc.init_params(cfg, seed=2, dtype=np.float64)).astype(np.float32))
    nf = dev(rng.normal(size=(n, 3)).astype(np.float32))
    ef = dev(rng.normal(size=(s.shape[0], 2)).astype(np.float32))
    tgt = dev(rng.normal(size=(n, 1)).astype(np.float32))
    mask = dev(np.arange(2, n, dtype=np.int32))
    graph = pkg.FeatureGraph(nf, ef, dev(s), dev(r))
    res = {}
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        model = pkg.Model(3, 2, 1, 2, 128, 2, compute_mode=mode)
        mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
        (gs,), loss = pkg.step_(mgn, graph, tgt, mask)
        res[mode] = (model.forward(graph, ps).cpu().numpy(), float(loss.cpu()), gs.cpu().numpy())
    a, b = res[pkg.COMPUTE_FP32], res[pkg.COMPUTE_BF16]
    assert rel(b[0], a[0]) < TOL_OUT and abs(a[1] - b[1]) < TOL_OUT * abs(a[1]) and rel(b[2], a[2]) < TOL_GRAD


def test_bf16_degenerate_graphs(pkg):
    """Edge cases of the tensor-core path against the fp32 path: no edges at all, and a graph whose tiles contain
    nodes without incoming edges (aggregation and receiver adjoint must write exact zeros for them)."""
    rng = np.random.default_rng(4)
    n = 200
    nf = dev(rng.normal(size=(n, 9)).astype(np.float32))
    tgt = dev(rng.normal(size=(n, 2)).astype(np.float32))
    mask = dev(np.arange(1, n + 1, dtype=np.int32))
    cfg = orc.ModelConfig(9, 3, 2, 128, 2, 2)
    ps = dev(orc.init_params(cfg, seed=6))
    cases = {
        "no_edges": (np.zeros(0, np.int32), np.zeros(0, np.int32)),
        "sparse": (rng.integers(1, n + 1, size=60).astype(np.int32), rng.integers(1, 40, size=60).astype(np.int32)),
    }
    for name, (s, r) in cases.items():
        ef = dev(rng.normal(size=(s.shape[0], 3)).astype(np.float32))
        graph = pkg.FeatureGraph(nf, ef, dev(s), dev(r))
        res = {}
        for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
            model = pkg.Model(9, 3, 2, 2, 128, 2, compute_mode=mode)
            mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
            (gs,), loss = pkg.step_(mgn, graph, tgt, mask)
            res[mode] = (model.forward(graph, ps).cpu().numpy(), float(loss.cpu()), gs.cpu().numpy())
            assert np.isfinite(res[mode][2]).all(), name
        a, b = res[pkg.COMPUTE_FP32], res[pkg.COMPUTE_BF16]
        assert rel(b[0], a[0]) < TOL_OUT and abs(a[1] - b[1]) < TOL_OUT * abs(a[1]) and rel(b[2], a[2]) < TOL_GRAD, name


@pytest.mark.parametrize("node_in,edge_in,out_dim", [(40, 7, 5), (64, 2, 16), (1, 1, 1)])
def test_bf16_feature_widths(pkg, node_in, edge_in, out_dim):
    """Raw feature widths up to the limits of the tensor-core path (<= 64 inputs, <= 16 outputs) against the fp32 path."""
    rng = np.random.default_rng(8)
    pos, cells, nt = orc.cylinder_flow_mesh(15, 11)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(node_in, edge_in, out_dim, 128, 2, 2)
    ps = dev(orc.init_params(cfg, seed=2))
    graph = pkg.FeatureGraph(dev(rng.normal(size=(N, node_in)).astype(np.float32)),
                             dev(rng.normal(size=(E, edge_in)).astype(np.float32)), dev(s), dev(r))
    tgt = dev(rng.normal(size=(N, out_dim)).astype(np.float32))
    mask = dev(orc.node_mask(nt, [0, 5]))
    res = {}
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        model = pkg.Model(node_in, edge_in, out_dim, 2, 128, 2, compute_mode=mode)
        mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
        (gs,), loss = pkg.step_(mgn, graph, tgt, mask)
        out = model.forward(graph, ps, training=True)
        dps, dnf = model.backward(graph, ps, torch.ones_like(out) / out.numel(), want_dnf=True)
        res[mode] = (out.cpu().numpy(), float(loss.cpu()), gs.cpu().numpy(), dnf.cpu().numpy())
    a, b = res[pkg.COMPUTE_FP32], res[pkg.COMPUTE_BF16]
    assert rel(b[0], a[0]) < TOL_OUT and abs(a[1] - b[1]) < TOL_OUT * abs(a[1])
    assert rel(b[2], a[2]) < TOL_GRAD and rel(b[3], a[3]) < 0.2
