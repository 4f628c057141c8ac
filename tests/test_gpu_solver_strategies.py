"""Solver-based training strategies on the GPU (src/strategies.jl SolverTraining :229-286, MultipleShooting :310-386
over ode_func_train, src/solve.jl:101-115) through the mirrored API (init_train_step / train_step), against the
SEQUENTIAL fp64 oracle (oracle/mgn_oracle_solver.py); the new C-ABI kernels against numpy; BASELINE configs[3]
(100k-node chain, multiple shooting) bf16 against fp32 mode.

Tolerances (relative L2 unless stated), fp32 mode / bf16 tensor-core mode against the fp64 oracle:
  predictions (state change over an interval)   1e-3          (same as the rollout tests)
  loss                                          1e-4 / 2e-2   (observed 3e-6 / 5e-3)
  parameter gradient                            2e-3 / 0.15   (observed 3e-4 / 4.5e-2: up to 12 pulled-back RHS evaluations)
  parameter gradient, one Tsit5 step            1e-2          (observed 3e-3, see below)

Tsit5 is only compared over ONE step per interval: the exact gradient of a multi-step fixed-step Tsit5 solve of a
ReLU network is ill-conditioned in single precision - the tableau's large cancelling coefficients (a_52 = -11.7,
a_62 = -12.9, b_5 = -3.3; sum_i b_i a_i3 = 0.3 from terms of size 25) multiply the stage-to-stage differences of the
piecewise-constant Jacobian, about two digits per step.  The numpy oracle shows the same loss of digits between its own
fp32 and fp64 runs (2e-6, 2e-4, 5e-2 after 1, 2, 4 steps; RK4 and Euler stay at 1e-7), so it is a property of the
method, not of the kernels; multi-step stage bookkeeping is covered by RK4 here and by Tsit5 in fp64 on the CPU
(tests/test_shooting_host.py).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

import mgn_oracle as orc
import mgn_oracle_solver as sol
from test_gpu_callers import _setup, dev, rel

pytestmark = pytest.mark.gpu

_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "solver_parity.jsonl")


def _log(**kw):
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _frozen_problem(pkg, mode, T=9, mps=3):
    """CylinderFlow-shaped 13x9 mesh, one accumulation of every online normaliser on both sides, then frozen."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=T, mps=mps, mode=mode)
    x0 = data_h["velocity"][0]
    for n_g, n_o, x in ((mgn.n_norm["velocity"], o["n_norm"]["velocity"], x0),
                        (mgn.e_norm, o["e_norm"], o["ef"]),
                        (mgn.o_norm["velocity"], o["o_norm"]["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x)); n_o(x)
        n_g.max_acc = 0.0; n_o.max_acc = np.float32(0)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    inflow_h = np.repeat((o["nt"] == 1)[:, None], 2, axis=1)
    data["node_type"] = dev(o["nt"].reshape(1, -1, 1).astype(np.int32))
    t = (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders, receivers, 1, None, dev(vm_h))
    rhs_o = sol.Rhs(o["cfg"], o["ps"], o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"], [2], {},
                    o["onehot"], o["ef"], o["s"], o["r"], vm_h, inflow_h, data_h["velocity"], 0.01, np.float64)
    return t, rhs_o, o, mgn


@pytest.mark.parametrize("mode,solver,n_sub,isz,tols", [
    (0, "euler", 1, 4, (1e-3, 1e-4, 2e-3)), (0, "euler", 2, 4, (1e-3, 1e-4, 2e-3)), (0, "rk4", 1, 4, (1e-3, 1e-4, 2e-3)),
    (0, "tsit5", 1, 2, (1e-3, 1e-4, 1e-2)), (1, "euler", 1, 4, (6e-2, 2e-2, 0.15)), (1, "rk4", 1, 4, (6e-2, 2e-2, 0.15))])
def test_multiple_shooting_step_matches_sequential_oracle(pkg, mode, solver, n_sub, isz, tols):
    """Intervals of 4, 4 and 2 observations (7 intervals of 2 for Tsit5) solved in lock-step on one block-diagonal
    graph vs the oracle's interval-by-interval solves: loss (incl. the continuity terms) and the parameter gradient."""
    t, rhs_o, o, mgn = _frozen_problem(pkg, mode)
    strat = pkg.MultipleShooting(0.0, 0.01, 0.07, solver, interval_size=isz, continuity_term=100, adaptive=False,
                                 dt=0.01 / n_sub)
    assert pkg.get_delta(strat, 9) == 1
    tt = pkg.init_train_step(strat, t, None)
    (gs,), loss = pkg.train_step(strat, tt)
    g_o, loss_o, preds_o = sol.train_step_multiple_shooting(rhs_o, 0.0, 0.01, 0.07, isz, 100, solver, n_sub)
    e_loss = abs(float(loss.cpu()) - loss_o) / abs(loss_o)
    e_g = rel(gs.cpu().numpy(), g_o)
    _log(test="multiple_shooting", mode=mode, solver=solver, n_sub=n_sub, interval_size=isz, loss=loss_o, e_loss=e_loss, e_grad=e_g)
    assert e_loss < tols[1] and e_g < tols[2]
    # the step is deterministic: a second evaluation is bitwise identical
    (gs2,), loss2 = pkg.train_step(strat, tt)
    assert torch.equal(gs, gs2) and torch.equal(loss, loss2)


@pytest.mark.parametrize("mode,solver,tstop,tols", [(0, "euler", 0.04, (1e-4, 2e-3)), (0, "rk4", 0.04, (1e-4, 2e-3)),
                                                     (0, "tsit5", 0.01, (1e-4, 1e-2)), (1, "euler", 0.04, (2e-2, 0.15))])
def test_solver_training_step_matches_oracle(pkg, mode, solver, tstop, tols):
    t, rhs_o, o, mgn = _frozen_problem(pkg, mode, T=6)
    strat = pkg.SolverTraining(0.0, 0.01, tstop, solver)
    (gs,), loss = pkg.train_step(strat, pkg.init_train_step(strat, t, None))
    g_o, loss_o, pred_o = sol.train_step_solver_training(rhs_o, o["n_norm"], ["velocity"], [2], 0.0, 0.01, tstop, solver)
    e_loss = abs(float(loss.cpu()) - loss_o) / abs(loss_o)
    e_g = rel(gs.cpu().numpy(), g_o)
    _log(test="solver_training", mode=mode, solver=solver, loss=loss_o, e_loss=e_loss, e_grad=e_g)
    assert e_loss < tols[0] and e_g < tols[1]


def test_lockstep_predictions_and_stage_workspace_modes(pkg):
    """Predictions of every interval against the oracle (fp32 mode, Tsit5 - the forward solve is well conditioned), and
    the two stage-workspace modes (activations of the 6 stages side by side / one stage at a time) give bitwise
    identical gradients."""
    t, rhs_o, o, mgn = _frozen_problem(pkg, 0)
    mgn_, data, inputs, fields, meta, tf, target_dict, node_type, ef, senders, receivers, vm, u0, gt = \
        pkg.init_train_step(pkg.SolverTraining(0.0, 0.01, 0.07, "tsit5"), t, None)
    inflow = (data["node_type"][0].reshape(-1) == 1)[:, None].repeat(1, 2)
    res = []
    for slots in (True, False):
        alg = pkg.DeviceAlgebra()
        rhs = pkg.DeviceRhs(mgn, mgn.ps, fields, tf, target_dict, inputs, node_type, ef, senders, receivers, vm, inflow,
                            gt, 3, alg)
        res.append(pkg.multiple_shooting_step(rhs, alg, mgn.ps, gt, vm, 0.0, 0.01, 0.07, 4, 100, "tsit5", 1,
                                              stage_slots=slots))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    _, _, preds_o = sol.train_step_multiple_shooting(rhs_o, 0.0, 0.01, 0.07, 4, 100, "tsit5", 1)
    assert [p.shape[0] for p in res[0][2]] == [4, 4, 2]
    for p, q in zip(res[0][2], preds_o):
        assert np.array_equal(p[0].cpu().numpy(), q[0].astype(np.float32))          # u0 is data
        assert rel(p[-1].cpu().numpy() - q[0], q[-1] - q[0]) < 1e-3


def test_interval_sharding_on_device(pkg):
    """world = 2 logical ranks: loss and gradient of the two shards sum to the unsharded step (fp32 summation order
    differs: 1e-5)."""
    t, rhs_o, o, mgn = _frozen_problem(pkg, 0)
    kw = dict(interval_size=3, continuity_term=10, adaptive=False)
    whole = pkg.MultipleShooting(0.0, 0.01, 0.08, "euler", **kw)
    (g,), loss = pkg.train_step(whole, pkg.init_train_step(whole, t, None))
    parts = []
    for r in range(2):
        s = pkg.MultipleShooting(0.0, 0.01, 0.08, "euler", rank=r, world=2, **kw)
        parts.append(pkg.train_step(s, pkg.init_train_step(s, t, None)))
    g_sum = parts[0][0][0] + parts[1][0][0]
    l_sum = float(parts[0][1].cpu()) + float(parts[1][1].cpu())
    assert abs(l_sum - float(loss.cpu())) < 1e-5 * abs(float(loss.cpu()))
    assert rel(g_sum.cpu().numpy(), g.cpu().numpy()) < 1e-5


def test_two_target_fields_and_an_input_field(pkg):
    """Column bookkeeping of the state: two target fields (velocity: 2, pressure: 1 -> S = 3) with a non-target input
    field between them in `fields`, each with its own normaliser kind; MultipleShooting (RK4, fp32 mode) against the
    sequential fp64 oracle."""
    rng = np.random.default_rng(21)
    pos, cells, nt = orc.cylinder_flow_mesh(11, 8)
    N, T = pos.shape[0], 7
    vel = orc.synthetic_velocity(pos, T, seed=3)
    prs = (0.5 * np.sin(3 * pos[None, :, :1] + 0.2 * np.arange(T)[:, None, None]) + 0.05 * rng.normal(size=(T, N, 1))).astype(np.float32)
    load = np.cos(5 * pos[:, 1:2]).astype(np.float32)[None].repeat(T, 0)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None]}
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}, "pressure": {"dim": 1}, "load": {"dim": 1}},
            "target_features": ["velocity", "pressure"]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0)
    fields = ["velocity", "load", "pressure"]
    model, ps, st = pkg.build_model(2 + 1 + 1 + 7, 2, 3, 2, 128, 2, compute_mode=pkg.COMPUTE_FP32, seed=9)
    n_g = {"velocity": pkg.NormaliserOnline(2), "load": pkg.NormaliserOfflineMinMax(-1.0, 1.0),
           "pressure": pkg.NormaliserOfflineMeanStd(0.1, 0.4), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)}
    o_g = {"velocity": pkg.NormaliserOnline(2), "pressure": pkg.NormaliserOfflineMeanStd(0.0, 5.0)}
    e_g = pkg.NormaliserOnline(3)
    n_o = {"velocity": orc.NormaliserOnline(2), "load": orc.NormaliserOfflineMinMax(-1.0, 1.0),
           "pressure": orc.NormaliserOfflineMeanStd(0.1, 0.4), "node_type": orc.NormaliserOfflineMinMax(0.0, 1.0)}
    o_o = {"velocity": orc.NormaliserOnline(2), "pressure": orc.NormaliserOfflineMeanStd(0.0, 5.0)}
    e_o = orc.NormaliserOnline(3)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    ef_h = orc.edge_features(pos, s, r)
    for g, o, x in ((n_g["velocity"], n_o["velocity"], vel[0]), (e_g, e_o, ef_h),
                    (o_g["velocity"], o_o["velocity"], (vel[1] - vel[0]) / np.float32(0.05))):
        g(dev(x)); o(x)
    mgn = pkg.GraphNetwork(model, ps, st, e_g, n_g, o_g)
    vm_h = orc.val_mask(nt, [0, 5], 3)
    data = {"velocity": dev(vel), "pressure": dev(prs), "load": dev(load),
            "node_type": dev(nt.reshape(1, -1, 1).astype(np.int32))}
    t = (mgn, data, meta, fields, ["velocity", "pressure"], node_type, ef, senders, receivers, 1, None, dev(vm_h))
    strat = pkg.MultipleShooting(0.0, 0.01, 0.06, "rk4", interval_size=4, continuity_term=50)
    tt = pkg.init_train_step(strat, t, None)
    assert tt[13].shape == (T, N, 3) and tt[6] == {"velocity": 2, "pressure": 1}          # gt = vcat(target features)
    (gs,), loss = pkg.train_step(strat, tt)
    cfg = orc.ModelConfig(11, 3, 3, 128, 2, 2)
    rhs_o = sol.Rhs(cfg, ps.cpu().numpy(), n_o, e_o, o_o, fields, ["velocity", "pressure"], [2, 1], {"load": load[0]},
                    orc.one_hot(nt, 7, 1), ef_h, s, r, vm_h, np.repeat((nt == 1)[:, None], 3, axis=1),
                    np.concatenate([vel, prs], axis=2), 0.01, np.float64)
    g_o, loss_o, _ = sol.train_step_multiple_shooting(rhs_o, 0.0, 0.01, 0.06, 4, 50, "rk4", 1)
    e_loss, e_g_ = abs(float(loss.cpu()) - loss_o) / abs(loss_o), rel(gs.cpu().numpy(), g_o)
    _log(test="two_target_fields", loss=loss_o, e_loss=e_loss, e_grad=e_g_)
    assert e_loss < 1e-4 and e_g_ < 2e-3


def test_chain_100k_multiple_shooting_bf16_vs_fp32(pkg):
    """BASELINE configs[3]: 1-D chain of 100 000 nodes (src/dataset.jl:379-382 through parse_edges) with a target
    field `u` (dim 1) and a non-target input field `load`, MultipleShooting (3 intervals in lock-step = a 300k-node
    block-diagonal graph, Euler, 2 MP steps to keep it short).  The oracle cannot run this size in seconds, so the
    tensor-core mode is compared with the library's fp32 mode (which the small cases pin to the oracle)."""
    n, T = 100_000, 7
    rng = np.random.default_rng(11)
    pos, edges, nt = orc.chain_mesh(n)
    xs = pos[:, 0]
    u = np.stack([np.sin(2 * np.pi * (xs - 0.05 * k)) + 0.05 * rng.normal(size=n) for k in range(T)])[:, :, None]
    load = np.cos(4 * np.pi * xs)[None, :, None].repeat(T, 0)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "edges": edges}
    meta = {"dt": 0.01, "features": {"u": {"dim": 1}, "load": {"dim": 1}}, "target_features": ["u"]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0)
    assert ef.shape == (2 * (n - 1), 2)
    data = {"u": dev(u.astype(np.float32)), "load": dev(load.astype(np.float32)),
            "node_type": dev(nt.reshape(1, -1, 1).astype(np.int32))}
    vm = dev(orc.val_mask(nt, [0], 1))
    res = {}
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        model, ps, st = pkg.build_model(1 + 1 + 7, 1, 1, 2, 128, 2, compute_mode=mode, seed=5)
        mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOfflineMeanStd(0.0, 1e-5),
                               {"u": pkg.NormaliserOfflineMeanStd(0.0, 0.7), "load": pkg.NormaliserOfflineMinMax(-1.0, 1.0),
                                "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                               {"u": pkg.NormaliserOfflineMeanStd(0.0, 3.0)})
        strat = pkg.MultipleShooting(0.0, 0.01, 0.06, "euler", interval_size=3, continuity_term=100)
        t = (mgn, data, meta, ["u", "load"], ["u"], node_type, ef, senders, receivers, 1, None, vm)
        (gs,), loss = pkg.train_step(strat, pkg.init_train_step(strat, t, None))
        assert torch.isfinite(gs).all() and float(gs.abs().max().cpu()) > 0
        res[mode] = (gs.cpu().numpy(), float(loss.cpu()))
    a, b = res[pkg.COMPUTE_FP32], res[pkg.COMPUTE_BF16]
    e_loss, e_g = abs(a[1] - b[1]) / abs(a[1]), rel(b[0], a[0])
    _log(test="chain_100k", loss=a[1], e_loss=e_loss, e_grad=e_g)
    assert e_loss < 3e-2 and e_g < 0.1


# ---------------------------------------------------------------------------------------------------------------
# The new C-ABI kernels against numpy
# ---------------------------------------------------------------------------------------------------------------


def test_lincomb_is_bit_exact_against_sequential_float32(pkg):
    rng = np.random.default_rng(0)
    n = 100_003
    x = rng.normal(size=n).astype(np.float32)
    ks = [rng.normal(size=n).astype(np.float32) for _ in range(11)]
    cs = [np.float32(c) for c in rng.normal(size=11)]
    cs[3] = np.float32(0)                                     # zero coefficients are skipped
    alg = pkg.DeviceAlgebra()
    for m in (1, 6, 8, 11):                                   # 11 terms: two launches of <= 8 terms
        want = x.copy()
        for k, c in zip(ks[:m], cs[:m]):
            if c != 0:
                want = want + c * k
        got = alg.lincomb(dev(x), [dev(k) for k in ks[:m]], cs[:m])
        assert np.array_equal(got.cpu().numpy(), want)
    acc = dev(x)                                              # in place, and x == None (zero start)
    alg.lincomb(acc, [dev(ks[0])], [1.0], out=acc)
    assert np.array_equal(acc.cpu().numpy(), x + ks[0])
    assert np.array_equal(alg.lincomb(None, [dev(ks[1])], [cs[1]]).cpu().numpy(), cs[1] * ks[1])


def test_overwrite_mul_and_strided_normalisers(pkg):
    rng = np.random.default_rng(1)
    x, src = rng.normal(size=(501, 3)).astype(np.float32), rng.normal(size=(501, 3)).astype(np.float32)
    mask = rng.random(size=(501, 3)) < 0.3
    alg = pkg.DeviceAlgebra()
    m8 = dev(mask.astype(np.uint8))
    assert np.array_equal(alg.overwrite(dev(x), dev(src), m8).cpu().numpy(), np.where(mask, src, x))
    assert np.array_equal(alg.overwrite(dev(x), None, m8).cpu().numpy(), np.where(mask, np.float32(0), x))
    assert np.array_equal(alg.mul(dev(x), dev(src)).cpu().numpy(), x * src)
    # strided maps: columns 1..2 of x -> columns 2..3 of a 5-wide output, all four modes, three normaliser kinds
    on_g, on_o = pkg.NormaliserOnline(2), orc.NormaliserOnline(2)
    on_g(dev(x[:, 1:3].copy())); on_o(x[:, 1:3])
    sd, mu = on_o.std(), on_o.mean()
    cases = [(on_g, lambda v: (v - mu) / sd, lambda v: v * sd + mu, 1 / sd, sd),
             (pkg.NormaliserOfflineMeanStd(0.5, 2.0), lambda v: (v - 0.5) / 2.0, lambda v: v * 2.0 + 0.5, 0.5, 2.0),
             (pkg.NormaliserOfflineMinMax(-1.0, 3.0), lambda v: (v + 1.0) / 4.0, lambda v: v * 4.0 - 1.0, 0.25, 4.0)]
    state_before = on_g.state.clone()
    for norm, fwd, inv, jf, ji in cases:
        for mode, fn in ((pkg.NORM_FORWARD, fwd), (pkg.NORM_INVERSE, inv), (pkg.NORM_FORWARD_VJP, lambda v: v * jf),
                         (pkg.NORM_INVERSE_VJP, lambda v: v * ji)):
            out = torch.full((501, 5), 7.0, device="cuda")
            norm.apply_ld(dev(x), 1, 2, mode, out, 2)
            o = out.cpu().numpy()
            assert np.allclose(o[:, 2:4], fn(x[:, 1:3].astype(np.float64)), rtol=2e-6, atol=1e-6)
            assert (o[:, :2] == 7).all() and (o[:, 4] == 7).all()
    assert torch.equal(on_g.state, state_before)              # apply_ld never accumulates


def test_shooting_losses_against_numpy(pkg):
    rng = np.random.default_rng(2)
    T, N, S = 5, 1203, 2
    pred, gt = rng.normal(size=(T, N, S)).astype(np.float32), rng.normal(size=(T, N, S)).astype(np.float32)
    vm = (rng.random(size=(N, 1)) < 0.8).astype(np.float32).repeat(S, 1)
    alg = pkg.DeviceAlgebra()
    loss = torch.full((1,), 3.0, device="cuda")
    dpred = torch.empty((T, N, S), device="cuda")
    w = 1.0 / (T * N * S)
    alg.mse(dev(pred), dev(gt), dev(vm), w, False, loss, dpred)
    d = gt.astype(np.float64) - pred
    assert abs(float(loss.cpu()) - (d * d * vm).mean()) < 1e-6
    assert np.allclose(dpred.cpu().numpy(), -2 * w * d * vm, rtol=1e-6, atol=1e-12)
    alg.mse(dev(pred), dev(gt), dev(vm), w, True, loss, dpred)            # accumulates
    assert abs(float(loss.cpu()) - 2 * (d * d * vm).mean()) < 2e-6
    a, b = pred[0], gt[0].copy()
    b[:7] = a[:7]                                                          # sign(0) = 0
    da = torch.ones((N, S), device="cuda")
    loss.zero_()
    alg.continuity(dev(a), dev(b), 100.0, loss, da)
    assert abs(float(loss.cpu()) - 100 * np.abs(a.astype(np.float64) - b).sum()) < 1e-5 * 100 * np.abs(a - b).sum()
    assert np.array_equal(da.cpu().numpy(), 1 + 100 * np.sign(a - b))
    l1 = loss.clone()
    loss.zero_(); da.fill_(1.0)
    alg.continuity(dev(a), dev(b), 100.0, loss, da)
    assert torch.equal(loss, l1)                                           # fixed summation order


def test_solver_abi_argument_errors(pkg):
    lib = pkg.load()
    x = torch.zeros(16, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    arr_k = (C.c_void_p * 9)(*[x.data_ptr()] * 9)
    arr_c = (C.c_float * 9)(*[1.0] * 9)
    assert lib.mgn_ode_lincomb(p(x), arr_k, arr_c, 9, 16, p(x), None) == 1          # MGN_ERR_INVALID: > 8 terms
    assert lib.mgn_ode_lincomb(p(x), None, None, 1, 16, p(x), None) == 1
    assert lib.mgn_masked_overwrite(p(x), None, None, 16, p(x), None) == 1          # null mask
    assert lib.mgn_norm_online_apply_ld(p(x), 2, 1, 4, 2, p(x), 1e-8, 0, p(x), 4, 0, None) == 1   # col_x + F > ld_x
    assert lib.mgn_norm_online_apply_ld(p(x), 2, 0, 4, 2, p(x), 1e-8, 7, p(x), 4, 0, None) == 1   # unknown mode
    assert lib.mgn_shooting_mse(p(x), p(x), p(x), 0, 4, 1.0, 0, p(x), p(x), None) == 1
    with pytest.raises(pkg.MgnError):
        t, rhs_o, o, mgn = _frozen_problem(pkg, 0)
        tt = pkg.init_train_step(pkg.SolverTraining(0.0, 0.01, 0.02, "euler"), t, None)
        rhs = pkg.DeviceRhs(mgn, mgn.ps, tt[3], tt[5], tt[6], tt[2], tt[7], tt[8], tt[9], tt[10], tt[11], None, tt[13], 1)
        rhs.backward(torch.zeros((tt[7].shape[0], 2), device="cuda"))               # no training forward saved
