"""build_graph (src/graph.jl:75-97) and `inverse_data(o_norm, out) .* val_mask` (src/solve.jl:205-218)
evaluated INSIDE the model kernels (mgn_forward_fused / mgn_backward_fused) against the operation-by-operation path
and the oracle; all online-normaliser updates of a step in two launches (mgn_norm_online_update_multi)."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc
from test_gpu_callers import _setup, dev, rel

pytestmark = pytest.mark.gpu


def _rhs_setup(pkg, mode, freeze):
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=6, mode=mode)
    x0 = data_h["velocity"][0]
    for n_g, n_o, x in ((mgn.n_norm["velocity"], o["n_norm"]["velocity"], x0), (mgn.e_norm, o["e_norm"], o["ef"]),
                        (mgn.o_norm["velocity"], o["o_norm"]["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x)); n_o(x)
        if freeze:
            n_g.max_acc = 0.0; n_o.max_acc = np.float32(0)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    p = (mgn, mgn.ps, {}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders, receivers, dev(vm_h))
    return data_h, mgn, o, x0, vm_h, p


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_rhs_equals_the_unfused_sequence(pkg, mode):
    """Same expressions, evaluated where the operand is staged / in the epilogue: the fused RHS reproduces the unfused
    one to fp32 rounding (bitwise when the compiler contracts both the same way) and the oracle to the mode's tolerance."""
    data_h, mgn, o, x0, vm_h, p = _rhs_setup(pkg, mode, freeze=True)
    a = pkg.ode_step(dev(x0), p, 0.0)
    b = pkg.ode_step_unfused(dev(x0), p, 0.0)
    torch.cuda.synchronize()
    print(f"[fused RHS mode {mode}] bitwise {torch.equal(a, b)} rel {rel(a.cpu().numpy(), b.cpu().numpy()):.2e}")
    assert rel(a.cpu().numpy(), b.cpu().numpy()) < (1e-6 if mode == 0 else 2e-3)
    rhs_o = orc.ode_step(o["cfg"], o["ps"], x0, o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"], [2],
                         {}, o["onehot"], o["ef"], o["s"], o["r"], vm_h, dtype=np.float64)
    assert rel(a.cpu().numpy(), rhs_o) < (1e-4 if mode == 0 else 3e-2)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_rhs_accumulates_like_build_graph(pkg, mode):
    """Normalisers that still accumulate: the fused RHS updates the statistics of the node-field and edge normalisers
    (two launches for all of them) exactly when build_graph's calls would, then normalises with the updated ones."""
    data_h, mgn_a, o, x0, vm_h, p_a = _rhs_setup(pkg, mode, freeze=False)
    _, mgn_b, _, _, _, p_b = _rhs_setup(pkg, mode, freeze=False)
    x1 = dev(data_h["velocity"][2])
    a = pkg.ode_step(x1, p_a, 0.0)
    b = pkg.ode_step_unfused(x1, p_b, 0.0)
    torch.cuda.synchronize()
    for na, nb in ((mgn_a.n_norm["velocity"], mgn_b.n_norm["velocity"]), (mgn_a.e_norm, mgn_b.e_norm)):
        assert float(na.state[-1].cpu()) == 2.0 and float(nb.state[-1].cpu()) == 2.0       # two accumulations each
        assert rel(na.state.cpu().numpy(), nb.state.cpu().numpy()) < 1e-6
    assert float(mgn_a.o_norm["velocity"].state[-1].cpu()) == 1.0                              # inverse_data never accumulates
    assert rel(a.cpu().numpy(), b.cpu().numpy()) < (1e-5 if mode == 0 else 3e-3)


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_pullback_matches_the_unfused_chain(pkg, mode):
    """What ZygoteVJP asks of ode_step (src/strategies.jl:183-194): d/d params and d/d x of <lambda, rhs(x)>.  Fused:
    mgn_backward_fused (mask, inverse-normaliser and normaliser transposes applied inside the kernels).  Unfused: scale the
    cotangent, mgn_backward, divide the state columns of d/d nf by the standard deviation."""
    data_h, mgn, o, x0, vm_h, p = _rhs_setup(pkg, mode, freeze=True)
    node_type, ef, senders, receivers = p[7], p[8], p[9], p[10]
    rng = np.random.default_rng(3)
    lam = dev(rng.normal(size=x0.shape).astype(np.float32))
    x = dev(x0)
    fg = pkg.FusedGraph([(mgn.n_norm["velocity"], x, 0, 2), (mgn.n_norm["node_type"], node_type, 0, 7)],
                        (mgn.e_norm, ef, 0, 3), senders, receivers, x.shape[0], [(mgn.o_norm["velocity"], 2)], dev(vm_h))
    out = pkg.forward_fused(mgn.model, fg, mgn.ps, training=True)
    dps, dx = pkg.backward_fused(mgn.model, fg, mgn.ps, lam, want_dx=True)
    # unfused chain
    graph = pkg.build_graph(mgn, {"velocity": x[None]}, ["velocity"], 1, node_type, ef, senders, receivers)
    mgn.model.forward(graph, mgn.ps, training=True, slot=1)
    st = mgn.o_norm["velocity"].state.cpu().numpy().astype(np.float64)
    sd_o = np.maximum(np.sqrt(st[2:4] / max(st[4], 1) - (st[0:2] / max(st[4], 1)) ** 2), 1e-8)
    sn = mgn.n_norm["velocity"].state.cpu().numpy().astype(np.float64)
    sd_n = np.maximum(np.sqrt(sn[2:4] / max(sn[4], 1) - (sn[0:2] / max(sn[4], 1)) ** 2), 1e-8)
    dout = lam * dev(vm_h) * dev(sd_o.astype(np.float32))[None]
    dps_u, dnf_u = mgn.model.backward(graph, mgn.ps, dout.contiguous(), want_dnf=True, slot=1)
    dx_u = dnf_u[:, :2] / dev(sd_n.astype(np.float32))[None]
    torch.cuda.synchronize()
    tol = 1e-5 if mode == 0 else 2e-3
    e_p, e_x = rel(dps.cpu().numpy(), dps_u.cpu().numpy()), rel(dx[:, :2].cpu().numpy(), dx_u.cpu().numpy())
    print(f"[fused pullback mode {mode}] d_params {e_p:.2e} d_x {e_x:.2e}")
    assert e_p < tol and e_x < tol
    assert rel(dx[:, 2:].cpu().numpy(), dnf_u[:, 2:].cpu().numpy()) < tol      # node_type block: MinMax(0, 1) is the identity


def test_multi_update_matches_single_updates_and_the_oracle(pkg):
    rng = np.random.default_rng(0)
    mats = [rng.normal(loc=3.0, scale=2.0, size=(5000, 5)).astype(np.float32), rng.normal(size=(777, 3)).astype(np.float32)]
    norms_a = [pkg.NormaliserOnline(2), pkg.NormaliserOnline(3), pkg.NormaliserOnline(3, max_acc=0.0)]
    norms_b = [pkg.NormaliserOnline(2), pkg.NormaliserOnline(3), pkg.NormaliserOnline(3, max_acc=0.0)]
    d0, d1 = dev(mats[0]), dev(mats[1])
    jobs = [(norms_a[0], d0, 1, 2), (norms_a[1], d1, 0, 3), (norms_a[2], d1, 0, 3)]          # a column block, a full matrix, a frozen one
    for _ in range(2):
        pkg.update_online(jobs)
        norms_b[0](d0[:, 1:3].contiguous()); norms_b[1](d1); norms_b[2](d1)
    torch.cuda.synchronize()
    ora = orc.NormaliserOnline(2)
    for _ in range(2):
        ora(mats[0][:, 1:3])
    want = np.concatenate([ora.acc_sum, ora.acc_sum_sq, [ora.acc_count, ora.num_acc]])
    assert rel(norms_a[0].state.cpu().numpy(), want) < 1e-6
    for a, b in zip(norms_a, norms_b):
        assert rel(a.state.cpu().numpy(), b.state.cpu().numpy()) < 1e-6 or float(b.state.abs().sum().cpu()) == 0.0
    assert float(norms_a[2].state.abs().sum().cpu()) == 0.0                                      # max_acc reached: untouched


def test_fused_io_argument_errors(pkg):
    data_h, mgn, o, x0, vm_h, p = _rhs_setup(pkg, 1, freeze=True)
    node_type, ef, senders, receivers = p[7], p[8], p[9], p[10]
    x = dev(x0)
    bad = pkg.FusedGraph([(mgn.n_norm["velocity"], x, 0, 2)], (mgn.e_norm, ef, 0, 3), senders, receivers, x.shape[0])
    with pytest.raises(pkg.MgnError) as e:                                                       # widths sum to 2, model needs 9
        pkg.forward_fused(mgn.model, bad, mgn.ps)
    assert e.value.code == 1
