"""The recalled GraphNetCore internals exposed as mgn_model_config switches: Dense layers per MLP,
LayerNorm parameter order, aggregation order.  CPU half: the flat parameter layout follows the switches exactly as the
oracle's does; GPU half: both arithmetic modes agree with the oracle under flipped switches."""
import numpy as np
import pytest

import mgn_oracle as orc


def _layout_of(specs):
    out = {}
    for s in specs:
        for l, (w, b, i, o) in enumerate(s.dense):
            out[f"{s.name}.dense{l + 1}.weight"] = (w, o, i)
            out[f"{s.name}.dense{l + 1}.bias"] = (b, o, 1)
        if s.ln is not None:
            out[f"{s.name}.layernorm.bias"] = (s.ln[0], s.out_dim, 1)
            out[f"{s.name}.layernorm.scale"] = (s.ln[1], s.out_dim, 1)
    return out


@pytest.mark.parametrize("dense_layers,scale_first", [(0, False), (0, True), (3, False), (2, True)])
def test_param_layout_follows_the_switches(pkg, dense_layers, scale_first):
    cfg = orc.ModelConfig(9, 3, 2, 128, 2, 2, dense_layers=dense_layers, ln_scale_first=scale_first)
    specs, P = orc.mlp_specs(cfg)
    model = pkg.Model(9, 3, 2, 2, 128, 2, compute_mode=pkg.COMPUTE_FP32, dense_layers=dense_layers,
                      ln_scale_first=scale_first)
    assert model.n_params == P
    want = _layout_of(specs)
    got = {name: (off, rows, cols) for name, off, rows, cols in model.param_layout()}
    assert got == want
    offs = [off for _, off, _, _ in model.param_layout()]
    assert offs == sorted(offs)                      # the table is listed in memory order


def test_bad_switch_values_fail_loudly(pkg):
    with pytest.raises(pkg.MgnError):
        pkg.Model(9, 3, 2, 2, 128, 2, dense_layers=1)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol,tol_dnf", [(0, 1e-5, 1e-5), (1, 6e-2, 0.3)])
@pytest.mark.parametrize("mps,nx,ny", [(3, 12, 9), (1, 5, 4), (5, 40, 21)])
def test_aggregate_post_residual_matches_oracle(pkg, mode, tol, tol_dnf, mps, nx, ny):
    """mgn_model_config::aggregate_post_residual = 1: agg = scatter(+, ef + m) instead of scatter(+, m) - forward
    (inference and training pass), loss, parameter gradient and d loss / d node features against the fp64 oracle, in both
    arithmetic modes; the switch must also CHANGE the result (it is not silently ignored)."""
    import torch
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
    if mode == 1 and nx * ny < 50:
        tol = 0.1          # a 20-node graph: single bf16 rounding flips are visible in the gradient (cf. smoke())
    rng = np.random.default_rng(7)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, 2, aggregate_post_residual=True)
    cfg0 = orc.ModelConfig(9, 3, 2, 128, mps, 2)
    P = orc.mlp_specs(cfg)[1]
    ps = (orc.init_params(cfg, seed=4, dtype=np.float64) + 0.02 * rng.normal(size=P)).astype(np.float32)
    nf, ef = rng.normal(size=(N, 9)).astype(np.float32), rng.normal(size=(E, 3)).astype(np.float32)
    tgt, mask = rng.normal(size=(N, 2)).astype(np.float32), orc.node_mask(nt, [0, 5])
    g_o, loss_o, out_o, dnf_o = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    out_pre = orc.model_forward(cfg0, ps.astype(np.float64), nf, ef, s, r, dtype=np.float64)
    assert rel(out_pre, out_o) > 0.1     # the two readings of the block are different functions
    model = pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=mode, aggregate_post_residual=True)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    out_inf = model.forward(graph, dev(ps), training=False)
    out_tr = model.forward(graph, dev(ps), training=True)
    assert torch.equal(out_inf, out_tr)
    assert rel(out_tr.cpu().numpy(), out_o) < max(tol, 5e-6)
    mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
    (gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
    assert abs(float(loss.cpu()) - loss_o) < max(tol, 1e-5) * abs(loss_o)
    assert rel(gs.cpu().numpy(), g_o) < tol
    # d loss / d node features through the same pullback
    _, dout_o = orc.loss_and_dout(out_o, tgt.astype(np.float64), mask)
    model.forward(graph, dev(ps), training=True)
    dps, dnf = model.backward(graph, dev(ps), dev(dout_o.astype(np.float32)), want_dnf=True)
    assert rel(dps.cpu().numpy(), g_o) < tol
    assert rel(dnf.cpu().numpy(), dnf_o) < tol_dnf


@pytest.mark.gpu
@pytest.mark.parametrize("mode,tol", [(0, 1e-5), (1, 6e-2)])
@pytest.mark.parametrize("dense_layers,scale_first", [(3, True), (0, True)])
def test_step_under_flipped_switches_matches_oracle(pkg, mode, tol, dense_layers, scale_first):
    import torch
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))
    rng = np.random.default_rng(2)
    pos, cells, nt = orc.cylinder_flow_mesh(12, 9)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, 3, 2, dense_layers=dense_layers, ln_scale_first=scale_first)
    P = orc.mlp_specs(cfg)[1]
    ps = (orc.init_params(cfg, seed=4, dtype=np.float64) + 0.05 * rng.normal(size=P)).astype(np.float32)  # LN scale / bias differ
    nf, ef = rng.normal(size=(N, 9)).astype(np.float32), rng.normal(size=(E, 3)).astype(np.float32)
    tgt, mask = rng.normal(size=(N, 2)).astype(np.float32), orc.node_mask(nt, [0, 5])
    g_o, loss_o, out_o, _ = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    model = pkg.Model(9, 3, 2, 3, 128, 2, compute_mode=mode, dense_layers=dense_layers, ln_scale_first=scale_first)
    mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
    (gs,), loss = pkg.step_(mgn, pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r)), dev(tgt), dev(mask))
    assert abs(float(loss.cpu()) - loss_o) < max(tol, 1e-5) * abs(loss_o)
    assert rel(gs.cpu().numpy(), g_o) < tol
