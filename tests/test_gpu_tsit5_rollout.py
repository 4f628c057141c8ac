"""BASELINE configs[2]: NeuralODE rollout through the src/solve.jl RHS with Tsit5 (src/solve.jl:58), 50 fixed steps,
both arithmetic modes, against the fp64 oracle's explicit Tsitouras 5(4) step (oracle/mgn_oracle_solver.py) with the
in-place inflow overwrite of src/solve.jl:151 at every stage.  Also: sub-stepping (dt dividing the save interval)."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc
import mgn_oracle_solver as ors
from test_gpu_callers import _setup, dev, rel

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _freeze(mgn, o, data_h):
    x0 = data_h["velocity"][0]
    for n_g, n_o, x in ((mgn.n_norm["velocity"], o["n_norm"]["velocity"], x0),
                        (mgn.e_norm, o["e_norm"], o["ef"]),
                        (mgn.o_norm["velocity"], o["o_norm"]["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x)); n_o(x)
        n_g.max_acc = 0.0; n_o.max_acc = np.float32(0)
    return x0


def _oracle_rollout(o, data_h, x0, saves, h, n_sub, solver, vm_h, inflow_h):
    tab = ors.TABLEAUS[solver]

    def f(x, t):   # ode_func_eval: overwrite the inflow nodes of the (stage) state IN PLACE, then ode_step
        idx = orc.inflow_index(t, saves[1] - saves[0])
        x[...] = np.where(inflow_h, data_h["velocity"][idx], x)
        return orc.ode_step(o["cfg"], o["ps"], x, o["n_norm"], o["e_norm"], o["o_norm"], ["velocity"], ["velocity"],
                            [2], {}, o["onehot"], o["ef"], o["s"], o["r"], vm_h, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    sol = [x.copy()]
    for i in range(len(saves) - 1):
        t = np.float32(saves[i])
        for j in range(n_sub):
            x = ors.rk_step(f, x, np.float32(t + np.float32(j) * np.float32(h)), h, tab)
            x = np.array(x)
        sol.append(x.copy())
    return sol


@pytest.mark.parametrize("mode,tol", [(0, 2e-3), (1, 8e-2)])
def test_50_step_tsit5_rollout_matches_oracle(pkg, mode, tol):
    """State change after 1, 10, 25, 50 Tsit5 steps (300 RHS evaluations): 2e-3 in fp32 mode, 8e-2 in the tensor-core
    mode (bf16 operands in every one of the 300 forwards)."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=52, mode=mode)
    x0 = _freeze(mgn, o, data_h)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    inflow_h = np.repeat((o["nt"] == 1)[:, None], 2, axis=1)
    saves = [np.float32(0.01) * i for i in range(51)]
    sol, ts = pkg.rollout(mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef,
                          senders, receivers, dev(vm_h), dev(inflow_h), data, 0.0, 0.5, 0.01, saves, solver="tsit5")
    sol_o = _oracle_rollout(o, data_h, x0, saves, 0.01, 1, "tsit5", vm_h, inflow_h)
    assert len(sol) == 51 and torch.isfinite(sol[-1]).all()
    errs = {i: rel(sol[i].cpu().numpy() - x0, sol_o[i] - x0) for i in (1, 10, 25, 50)}
    print(f"[tsit5 50 steps, mode {mode}] " + " ".join(f"{i}: {e:.2e}" for i, e in errs.items()))
    assert all(e < tol for e in errs.values()), errs


def test_rollout_sub_steps_when_dt_divides_the_save_interval(pkg):
    """dt = 0.005 with saves every 0.01 (solve(...; dt = dt, saveat = saves), src/solve.jl:62): two Euler steps per save,
    the inflow data indexed by each sub-step's own time."""
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=12, mode=0)
    x0 = _freeze(mgn, o, data_h)
    vm_h = orc.val_mask(o["nt"], [0, 5], 2)
    inflow_h = np.repeat((o["nt"] == 1)[:, None], 2, axis=1)
    saves = [np.float32(0.01) * i for i in range(6)]
    sol, ts = pkg.rollout(mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef,
                          senders, receivers, dev(vm_h), dev(inflow_h), data, 0.0, 0.05, 0.005, saves)
    sol_o = _oracle_rollout(o, data_h, x0, saves, 0.005, 2, "euler", vm_h, inflow_h)
    assert len(sol) == 6
    for i in (1, 3, 5):
        assert rel(sol[i].cpu().numpy() - x0, sol_o[i] - x0) < 1e-3, i
    with pytest.raises(pkg.MgnError):
        pkg.rollout(mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef,
                    senders, receivers, dev(vm_h), dev(inflow_h), data, 0.0, 0.05, 0.004, saves)
