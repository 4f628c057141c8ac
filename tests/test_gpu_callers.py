"""The callers either side of the hot path, through the mirrored reference API: create_base_graph /
build_graph (src/graph.jl), init_train_step / train_step (src/strategies.jl:395-422), the
optimiser update (src/MeshGraphNets.jl:374-378) and ode_step / rollout (src/solve.jl)."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _setup(pkg, nx=13, ny=9, T=6, mps=3, mode=0, seed=0):
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    vel = orc.synthetic_velocity(pos, T + 1, seed=seed)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None],
              "velocity": vel[:T], "target|velocity": vel[1:T + 1]}
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}, "node_type": {"data_min": 0, "data_max": 6}},
            "target_features": ["velocity"]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0)
    data = {"velocity": dev(data_h["velocity"]), "target|velocity": dev(data_h["target|velocity"])}
    model, ps, st = pkg.build_model(2 + 7, 2, 2, mps, 128, 2, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3), {"velocity": pkg.NormaliserOnline(2),
                           "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)}, {"velocity": pkg.NormaliserOnline(2)})
    # oracle twins
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, 2)
    o = {"cfg": cfg, "ps": ps.cpu().numpy().copy(), "e_norm": orc.NormaliserOnline(3),
         "n_norm": {"velocity": orc.NormaliserOnline(2), "node_type": orc.NormaliserOfflineMinMax(0.0, 1.0)},
         "o_norm": {"velocity": orc.NormaliserOnline(2)}}
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    o.update(s=s, r=r, onehot=orc.one_hot(nt, 7, 1), ef=orc.edge_features(pos, s, r), nt=nt)
    return data_h, data, meta, mgn, (node_type, senders, receivers, ef), o


def test_create_base_graph_bit_exact(pkg):
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg)
    assert np.array_equal(senders.cpu().numpy(), o["s"]) and np.array_equal(receivers.cpu().numpy(), o["r"])
    assert np.array_equal(node_type.cpu().numpy(), o["onehot"])
    assert np.array_equal(ef.cpu().numpy(), o["ef"])  # fp32 subtraction + widened norm: bit-exact


def test_derivative_training_steps_match_oracle(pkg):
    """3 derivative-training steps incl. online-normaliser accumulation and Adam.  fp32 mode,
    tolerance 5e-4 on each step's loss (which sees the previous steps' updates) and 5e-2 on the
    parameters' total change: Adam divides every gradient component by its own magnitude, so
    components with |g| near fp32 noise turn rounding differences into O(lr) differences."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg)
    mask_h = orc.node_mask(o["nt"], [0, 5])
    mask = dev(mask_h)
    opt = pkg.Adam(1e-4)
    state = opt.setup(mgn.ps)
    strat = pkg.DerivativeTraining()
    ps0 = o["ps"].copy()
    m = np.zeros_like(ps0); v = np.zeros_like(ps0)
    for dp in range(1, 4):
        t = pkg.init_train_step(strat, (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders,
                                        receivers, dp, mask, None))
        gs, loss = pkg.train_step(strat, t)
        for g in gs:
            state, mgn.ps = opt.update(state, mgn.ps, g)
        # oracle: strategies.jl:399-410 then graph.jl:75-97 then step! then Optimisers.update
        tgt = o["o_norm"]["velocity"]((data_h["target|velocity"][dp - 1] - data_h["velocity"][dp - 1]) / np.float32(0.01))
        nf, efn = orc.build_graph_features(o["n_norm"], o["e_norm"], {"velocity": data_h["velocity"][dp - 1]},
                                           ["velocity"], o["onehot"], o["ef"])
        assert rel(t[1].node_features.cpu().numpy(), nf) < 1e-4
        assert rel(t[1].edge_features.cpu().numpy(), efn) < 1e-4
        g_o, loss_o, _, _ = orc.step(o["cfg"], o["ps"].astype(np.float64), nf, efn, o["s"], o["r"], tgt, mask_h,
                                     dtype=np.float64)
        assert abs(float(loss.cpu()) - loss_o) < 5e-4 * abs(loss_o)
        o["ps"], m, v = orc.adam_update(o["ps"], g_o.astype(np.float32), m, v, dp, lr=1e-4)
    assert rel(mgn.ps.cpu().numpy() - ps0, o["ps"] - ps0) < 5e-2


def test_ode_step_and_euler_rollout_match_oracle(pkg):
    """src/solve.jl:188-219 RHS and a 10-step fixed-step Euler rollout (cylinder_flow.jl:79-84).
    fp32 mode; tolerance 1e-3 relative on the final state's change."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=12)
    # fix the normaliser statistics on both sides (one accumulation each), then freeze them
    x0 = data_h["velocity"][0]
    for n_g, n_o, x in ((mgn.n_norm["velocity"], o["n_norm"]["velocity"], x0),
                        (mgn.e_norm, o["e_norm"], o["ef"]),
                        (mgn.o_norm["velocity"], o["o_norm"]["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x)); n_o(x)
        n_g.max_acc = 0.0; n_o.max_acc = np.float32(0)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    inflow_h = np.repeat((o["nt"] == 1)[:, None], 2, axis=1)
    vm, inflow = dev(vm_h), dev(inflow_h)
    target_dict = {"velocity": 2}
    p = (mgn, mgn.ps, {}, ["velocity"], meta, ["velocity"], target_dict, node_type, ef, senders, receivers, vm)
    rhs = pkg.ode_step(dev(x0), p, 0.0)
    rhs_o = orc.ode_step(o["cfg"], o["ps"], x0, o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"], [2],
                         {}, o["onehot"], o["ef"], o["s"], o["r"], vm_h, dtype=np.float64)
    assert rel(rhs.cpu().numpy(), rhs_o) < 1e-4
    saves = [0.01 * i for i in range(11)]
    sol, ts = pkg.rollout(mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], target_dict, node_type, ef,
                          senders, receivers, vm, inflow, data, 0.0, 0.1, 0.01, saves)

    def inflow_fn(x, t):
        return np.where(inflow_h, data_h["velocity"][orc.inflow_index(t, saves[1] - saves[0])], x)
    sol_o = orc.rollout_euler(
        lambda x, t: orc.ode_step(o["cfg"], o["ps"], x, o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"],
                                  [2], {}, o["onehot"], o["ef"], o["s"], o["r"], vm_h, dtype=np.float64),
        x0, saves, 0.01, inflow_fn)
    assert len(sol) == 11
    for i in (1, 2, 5):
        assert rel(sol[i].cpu().numpy() - x0, sol_o[i] - x0) < 1e-3, i
    assert rel(sol[-1].cpu().numpy() - x0, sol_o[-1] - x0) < 1e-3


@pytest.mark.parametrize("mode,tol", [(0, 2e-3), (1, 6e-2)])
def test_50_step_euler_rollout_matches_oracle(pkg, mode, tol):
    """BASELINE configs[2] shape of the check: state after a 50-step fixed-step rollout through the src/solve.jl RHS
    mirror (inflow overwrite each step) against the fp64 oracle.  Tolerance on the accumulated state change: 2e-3 in
    fp32 mode, 6e-2 in the bf16 tensor-core mode (50 RHS evaluations of bf16 arithmetic)."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=52, mode=mode)
    x0 = data_h["velocity"][0]
    for n_g, n_o, x in ((mgn.n_norm["velocity"], o["n_norm"]["velocity"], x0),
                        (mgn.e_norm, o["e_norm"], o["ef"]),
                        (mgn.o_norm["velocity"], o["o_norm"]["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x)); n_o(x)
        n_g.max_acc = 0.0; n_o.max_acc = np.float32(0)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    inflow_h = np.repeat((o["nt"] == 1)[:, None], 2, axis=1)
    saves = [np.float32(0.01) * i for i in range(51)]
    sol, ts = pkg.rollout(mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef,
                          senders, receivers, dev(vm_h), dev(inflow_h), data, 0.0, 0.5, 0.01, saves)

    def inflow_fn(x, t):
        return np.where(inflow_h, data_h["velocity"][orc.inflow_index(t, saves[1] - saves[0])], x)
    sol_o = orc.rollout_euler(
        lambda x, t: orc.ode_step(o["cfg"], o["ps"], x, o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"],
                                  [2], {}, o["onehot"], o["ef"], o["s"], o["r"], vm_h, dtype=np.float64),
        x0, saves, 0.01, inflow_fn)
    assert len(sol) == 51 and torch.isfinite(sol[-1]).all()
    for i in (1, 10, 25, 50):
        assert rel(sol[i].cpu().numpy() - x0, sol_o[i] - x0) < tol, i


@pytest.mark.parametrize("solver", ["euler", "tsit5"])
def test_captured_rollout_replays_the_eager_rollout(pkg, solver):
    """the whole rollout (inflow overwrite, build_graph normalisers, model, inverse_data, val_mask and
    the Runge-Kutta combinations of every step) captured once as ONE CUDA graph; replays are bitwise identical to the
    eager rollout, also from a different initial state."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=12, mode=1)
    x0 = data_h["velocity"][0]
    for n_g, x in ((mgn.n_norm["velocity"], x0), (mgn.e_norm, o["ef"]),
                   (mgn.o_norm["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x))
        n_g.max_acc = 0.0
    vm = dev(orc.val_mask(o["nt"], [0, 5], 2))
    inflow = dev(np.repeat((o["nt"] == 1)[:, None], 2, axis=1))
    saves = [np.float32(0.01) * i for i in range(11)]
    args = (mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders,
            receivers, vm, inflow, data, 0.0, 0.1, 0.01, saves)
    sol, ts = pkg.rollout(*args, solver=solver)
    cap = pkg.CapturedRollout(*args, solver=solver)
    for _ in range(2):
        sol_g, ts_g = cap.replay()
        assert ts_g == ts and len(sol_g) == 11
        assert all(torch.equal(a, b) for a, b in zip(sol, sol_g))
    x1 = {"velocity": dev(data_h["velocity"][1])}
    sol2, _ = pkg.rollout(mgn, x1, *args[2:], solver=solver)
    sol2_g, _ = cap.replay(x1)
    assert all(torch.equal(a, b) for a, b in zip(sol2, sol2_g))
    assert not torch.equal(sol2_g[-1], sol[-1])


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n_buckets", [1, 4, 64])
def test_backward_dp_single_gpu_equals_step_then_adam(pkg, mode, n_buckets):
    """SURVEY 8f row 1 / src/MeshGraphNets.jl:374-378: mgn_backward_dp without a communicator = backward with the Adam
    update of every finished gradient bucket running on a side stream beside the rest of the backward pass.  The same
    kernels on the same values: gradient, parameters and moments are BITWISE those of step! followed by the update."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, mode=mode)
    mask = dev(orc.node_mask(o["nt"], [0, 5]))
    strat = pkg.DerivativeTraining()
    t = pkg.init_train_step(strat, (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders, receivers, 1,
                                    mask, None))
    _, graph, target, _ = t
    ps0 = mgn.ps.clone()
    opt = pkg.Adam(1e-3)
    st_a, st_b = opt.setup(mgn.ps), opt.setup(mgn.ps)
    losses = []
    for _ in range(2):                                  # two steps: the device-side step counter must advance once per step
        (g_a,), loss_a = pkg.step_(mgn, graph, target, mask)
        g_a = g_a.clone()
        opt.update(st_a, mgn.ps, g_a)
        losses.append(float(loss_a.cpu()))
    ps_a, mgn.ps = mgn.ps.clone(), ps0.clone()
    for i in range(2):
        (g_b,), loss_b = pkg.step_dp_(mgn, graph, target, mask, opt=opt, opt_state=st_b, comm=None, n_buckets=n_buckets)
        torch.cuda.synchronize()
        assert float(loss_b.cpu()) == losses[i]
    assert torch.equal(g_b, g_a)
    assert torch.equal(mgn.ps, ps_a)
    assert torch.equal(st_a["m"], st_b["m"]) and torch.equal(st_a["v"], st_b["v"])
    assert int(st_b["dev"][0].cpu()) == 2
