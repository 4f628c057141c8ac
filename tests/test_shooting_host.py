"""Host logic of the lock-step shooting engine (meshgraphnets.jl_b200/shooting.py) on CPU.

The engine is generic over two interfaces (`rhs`, `alg`) whose only PRODUCT implementations call libmgn_b200
(DeviceRhs / DeviceAlgebra - exercised on the GPU by tests/test_gpu_solver_strategies.py).  Here they are replaced
by test doubles built on the oracle, so that the engine's own bookkeeping - interval batching into one
block-diagonal state, per-interval times and inflow indices, Runge-Kutta stage scheduling, the checkpointed reverse
sweep with per-stage slots, loss assembly with the continuity terms, interval sharding over ranks - is compared with
the SEQUENTIAL oracle (oracle/mgn_oracle_solver.py, which solves interval after interval as the reference does).
The doubles live in tests/ only; the product has no CPU path."""
import numpy as np
import pytest
import torch

import mgn_oracle as orc
import mgn_oracle_solver as sol
from test_oracle_solver import _problem


class NumpyAlgebra:
    """Stand-in for DeviceAlgebra (fp64, CPU tensors)."""

    def lincomb(self, x, ks, coefs, out=None):
        acc = torch.zeros_like(ks[0]) if x is None else x.clone()
        for k, c in zip(ks, coefs):
            if c != 0.0:
                acc = acc + float(c) * k
        if out is None:
            return acc
        out.copy_(acc)
        return out

    def mse(self, pred, gt, vm, weight, accumulate, loss, dpred):
        d = gt - pred
        val = weight * (d * d * vm).sum()
        loss[0] = (loss[0] if accumulate else 0.0) + val
        dpred.copy_(-2.0 * weight * d * vm)

    def continuity(self, a, b, weight, loss, da):
        d = a - b
        loss[0] = loss[0] + weight * d.abs().sum()
        da.add_(weight * torch.sign(d))


class OracleRhs:
    """Stand-in for DeviceRhs: K intervals, each evaluated by the oracle right-hand side."""

    def __init__(self, rhs: sol.Rhs, K, N):
        self.rhs, self.K, self.N = rhs, K, N
        self._saved = {}
        self.n_evals = 0
        self.max_live_slots = 0

    def state_norm(self, x, mode, out=None):
        a = np.concatenate([np.broadcast_to(self.rhs.nn[f][0], (d,)) for f, d in zip(self.rhs.tf, self.rhs.td)])
        c = np.concatenate([np.broadcast_to(self.rhs.nn[f][1], (d,)) for f, d in zip(self.rhs.tf, self.rhs.td)])
        a, c = torch.from_numpy(a.copy()), torch.from_numpy(c.copy())
        return x * a + c if mode == 0 else x * a          # NORM_FORWARD / NORM_FORWARD_VJP

    def forward(self, x, data_idx, training=False, slot=0):
        self.n_evals += 1
        assert len(data_idx) == self.K
        xs = x.numpy().reshape(self.K, self.N, -1)
        out = np.stack([self.rhs(xs[k], 0.0, idx=data_idx[k]) for k in range(self.K)])
        if training:
            self._saved[slot] = (xs.copy(), list(data_idx))
            self.max_live_slots = max(self.max_live_slots, len(self._saved))
        return torch.from_numpy(out.reshape(self.K * self.N, -1))

    def backward(self, dy, slot=0):
        xs, idx = self._saved.pop(slot)
        dys = dy.numpy().reshape(self.K, self.N, -1)
        g, dx = 0.0, []
        for k in range(self.K):
            gk, dxk = self.rhs.vjp(xs[k], 0.0, dys[k], idx=idx[k])
            g = g + gk
            dx.append(dxk)
        return torch.from_numpy(g), torch.from_numpy(np.stack(dx).reshape(self.K * self.N, -1))


def _run(pkg_mod, make, p, N, owned=None, **kw):
    rhs = make(p)
    n_int = len(pkg_mod.shooting_ranges(len(pkg_mod.time_steps(kw["tstart"], kw["dt"], kw["tstop"])), kw["interval_size"]))
    K = n_int if owned is None else len(owned)
    double = OracleRhs(rhs, max(K, 1), N)
    gt, vm = torch.from_numpy(rhs.gt), torch.from_numpy(rhs.vm)
    g, loss, preds = pkg_mod.multiple_shooting_step(double, NumpyAlgebra(), torch.from_numpy(p), gt, vm, owned=owned, **kw)
    return g.numpy(), float(loss[0]), preds, double


@pytest.fixture(scope="module")
def shooting():
    import mgn_pkg
    return mgn_pkg.pkg


@pytest.mark.parametrize("solver,n_sub,slots", [("euler", 1, True), ("euler", 2, True), ("rk4", 1, True),
                                                ("tsit5", 1, True), ("tsit5", 1, False)])
def test_lockstep_engine_equals_sequential_oracle(shooting, solver, n_sub, slots):
    """4 intervals (the last one shorter) advanced as one batch == the oracle's interval-by-interval solves."""
    cfg, p, make, _, N = _problem(seed=2, T=9)
    kw = dict(tstart=0.0, dt=0.01, tstop=0.07, interval_size=4, continuity_term=100, solver=solver, n_sub=n_sub)
    g_ref, loss_ref, preds_ref = sol.train_step_multiple_shooting(make(p), **kw)
    assert [q.shape[0] for q in preds_ref] == [4, 4, 2]
    g, loss, preds, double = _run(shooting, make, p, N, stage_slots=slots, **kw)
    assert abs(loss - loss_ref) < 1e-11 * abs(loss_ref)
    for a, b in zip(preds, preds_ref):
        assert np.allclose(a.numpy(), b, rtol=1e-11, atol=1e-13)
    assert np.allclose(g, g_ref, rtol=1e-8, atol=1e-12 * np.abs(g_ref).max())
    s = len(shooting.RK_TABLEAUS[solver][2])
    assert double.max_live_slots == (s if slots and s > 1 else 1)      # stages side by side, or one at a time
    assert not double._saved                                            # every saved stage was consumed


def test_interval_sharding_sums_to_the_whole(shooting):
    """Ranks own intervals r, r+world, ...; loss and gradient summed over ranks == the unsharded step."""
    cfg, p, make, _, N = _problem(seed=3, T=9)
    kw = dict(tstart=0.0, dt=0.01, tstop=0.08, interval_size=3, continuity_term=10, solver="rk4", n_sub=1)
    g_all, loss_all, _, _ = _run(shooting, make, p, N, **kw)
    n_int = len(shooting.shooting_ranges(9, 3))
    for world in (2, 3, 5):
        parts = [_run(shooting, make, p, N, owned=shooting.shard_intervals(n_int, r, world), **kw) for r in range(world)]
        assert sorted(i for r in range(world) for i in shooting.shard_intervals(n_int, r, world)) == list(range(n_int))
        assert abs(sum(q[1] for q in parts) - loss_all) < 1e-11 * abs(loss_all)
        assert np.allclose(sum(q[0] for q in parts), g_all, rtol=1e-9, atol=1e-12 * np.abs(g_all).max())


def test_solver_training_engine_equals_oracle(shooting):
    cfg, p, make, n_norms, N = _problem(seed=4, T=5)
    rhs = make(p)
    g_ref, loss_ref, pred_ref = sol.train_step_solver_training(make(p), n_norms, ["velocity"], [2], 0.0, 0.01, 0.04,
                                                               solver="tsit5")
    double = OracleRhs(rhs, 1, N)
    g, loss, pred = shooting.solver_training_step(double, NumpyAlgebra(), torch.from_numpy(p), torch.from_numpy(rhs.gt),
                                                  torch.from_numpy(rhs.vm), 0.0, 0.01, 0.04, solver="tsit5")
    assert abs(float(loss[0]) - loss_ref) < 1e-11 * abs(loss_ref)
    assert np.allclose(pred.numpy(), pred_ref, rtol=1e-11, atol=1e-13)
    assert np.allclose(g.numpy(), g_ref, rtol=1e-8, atol=1e-12 * np.abs(g_ref).max())


def test_strategy_constructors_and_errors(shooting):
    s = shooting.MultipleShooting(0.0, 0.01, 0.49, "Tsit5", interval_size=6, adaptive=False, dt=0.005)
    assert (s.solver, s.n_sub, s.interval_size, s.continuity_term) == ("tsit5", 2, 6, 100)
    assert shooting.get_delta(s, 600) == 1 and shooting.get_delta(shooting.DerivativeTraining(), 600) == 599
    assert shooting.SolverTraining(0.0, 0.01, 0.49, "euler").n_sub == 1
    with pytest.raises(shooting.MgnError):
        shooting.SolverTraining(0.0, 0.01, 0.49, "rodas5")               # not an explicit fixed-step method here
    with pytest.raises(shooting.MgnError):
        shooting.SolverTraining(0.0, 0.01, 0.49, "tsit5", adaptive=True)
    with pytest.raises(shooting.MgnError):
        shooting.SolverTraining(0.0, 0.01, 0.49, "euler", dt=0.003)      # does not divide the save interval
    with pytest.raises(shooting.MgnError):
        shooting.MultipleShooting(0.0, 0.01, 0.49, "euler", interval_size=1)
    with pytest.raises(ValueError):                                      # more time points than data
        cfg, p, make, _, N = _problem(seed=2, T=4)
        _run(shooting, make, p, N, tstart=0.0, dt=0.01, tstop=0.07, interval_size=4)


def test_product_time_grid_and_interval_helpers_known_answers(shooting):
    """The product's own restatement of `tstart:dt:tstop`, the interval ranges (src/strategies.jl:344-347) and the
    inflow data index (src/solve.jl:106) against hand-derived values and against the oracle's."""
    ts = shooting.time_steps(0.0, 0.01, 0.49)
    assert ts.dtype == np.float32 and ts.shape == (50,) and ts[-1] == np.float32(0.49) and ts[7] == np.float32(0.07)
    assert np.array_equal(ts, sol.tsteps(0.0, 0.01, 0.49))
    assert shooting.time_steps(0.5, 0.25, 1.3).tolist() == [0.5, 0.75, 1.0, 1.25]          # stop not on the grid
    with pytest.raises(ValueError):
        shooting.time_steps(0.0, 0.0, 1.0)
    assert shooting.shooting_ranges(11, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)] == sol.shooting_ranges(11, 4)
    assert shooting.shooting_ranges(50, 6) == sol.shooting_ranges(50, 6) and len(shooting.shooting_ranges(50, 6)) == 10
    with pytest.raises(ValueError):
        shooting.shooting_ranges(11, 1)
    for n in (0, 1, 7, 40):
        for world in (1, 2, 3, 8):
            parts = [shooting.shard_intervals(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n)) and max(map(len, parts)) - min(map(len, parts)) <= 1
    # floor(Int, t / dt) + 1 in Float32, 0-based here, clamped to the data
    from meshgraphnets_jl_b200.shooting import inflow_index
    assert [inflow_index(np.float32(0.01) * k, 0.01, 100) for k in range(5)] == [0, 1, 2, 3, 4]
    assert inflow_index(np.float32(0.0198), 0.01, 100) == 1 and inflow_index(np.float32(0.07), 0.01, 5) == 4
    assert all(inflow_index(t, 0.01, 100) == orc.inflow_index(t, 0.01) for t in np.linspace(0, 0.9, 181, dtype=np.float32))
