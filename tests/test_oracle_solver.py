"""The solver-strategy oracle (oracle/mgn_oracle_solver.py): interval ranges and time grids against
hand-derived values of the Julia expressions (src/strategies.jl:346-347), the Runge-Kutta tableaus
against their order conditions and a known ODE, and the hand-written discrete adjoint against finite
differences in fp64 - so that the GPU solver strategies are compared with a verified gradient."""
import numpy as np
import pytest

import mgn_oracle as orc
import mgn_oracle_solver as sol


def test_shooting_ranges_known_answers():
    # tsteps = 0:0.01:0.1 (11 points), interval_size 4: 1:3:10 -> 1:4, 4:7, 7:10, 10:11
    assert sol.shooting_ranges(11, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    # interval_size 6: 1:5:10 -> 1:6, 6:11
    assert sol.shooting_ranges(11, 6) == [(0, 5), (5, 10)]
    # one interval covering everything
    assert sol.shooting_ranges(5, 5) == [(0, 4)]
    # consecutive intervals share their boundary point (the continuity term compares exactly those)
    rg = sol.shooting_ranges(50, 7)
    assert all(rg[i][1] == rg[i + 1][0] for i in range(len(rg) - 1)) and rg[-1][1] == 49


def test_tsteps_matches_julia_range():
    ts = sol.tsteps(0.0, 0.01, 0.49)
    assert ts.shape == (50,) and ts.dtype == np.float32
    assert ts[0] == 0 and ts[-1] == np.float32(0.49) and ts[3] == np.float32(0.03)
    assert sol.tsteps(0.0, 0.1, 0.35).shape == (4,)


@pytest.mark.parametrize("name,order", [("euler", 1), ("rk4", 4), ("tsit5", 5)])
def test_tableaus_satisfy_order_conditions(name, order):
    c, A, b = sol.TABLEAUS[name]
    b = np.asarray(b)
    cc = np.asarray((0.0,) + tuple(c))
    assert abs(b.sum() - 1) < 1e-12
    for row, ci in zip(A, c):
        assert abs(sum(row) - ci) < 1e-12          # row-sum condition
    for q in range(2, order + 1):
        assert abs((b * cc ** (q - 1)).sum() - 1.0 / q) < 1e-9   # quadrature conditions
    # convergence order on x' = -x, x(0) = 1
    f = lambda x, t: -x
    errs = []
    for n in (2, 4):   # coarse steps: the stage coefficients are rounded to Float32 (1e-8 floor)
        x = np.array([1.0])
        for i in range(n):
            x = sol.rk_step(f, x, i / n, 1.0 / n, sol.TABLEAUS[name])
        errs.append(abs(x[0] - np.exp(-1.0)))
    assert np.log2(errs[0] / errs[1]) > order - 0.35


def _problem(seed=0, T=7, D=8, mps=1, nx=4, ny=3, dtype=np.float64):
    rng = np.random.default_rng(seed)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N = pos.shape[0]
    cfg = orc.ModelConfig(node_in=2 + 7, edge_in=3, out_dim=2, latent=D, mps=mps)
    p = orc.init_params(cfg, dtype=np.float64) + 0.05 * rng.normal(size=orc.mlp_specs(cfg)[1])
    gt = orc.synthetic_velocity(pos, T, seed=seed).astype(np.float64)
    n_norms = {"velocity": orc.NormaliserOnline(2), "node_type": orc.NormaliserOfflineMinMax(0.0, 1.0)}
    o_norms = {"velocity": orc.NormaliserOnline(2)}
    e_norm = orc.NormaliserOnline(3)
    ef = orc.edge_features(pos, s, r)
    for t in range(3):
        n_norms["velocity"](gt[t].astype(np.float32))
        o_norms["velocity"](((gt[t + 1] - gt[t]) / 0.01).astype(np.float32))
    e_norm(ef)
    vm = orc.val_mask(nt, [0, 5], 2)
    inflow = np.repeat((nt == 1)[:, None], 2, axis=1)

    def make(params):
        return sol.Rhs(cfg, params, n_norms, e_norm, o_norms, ["velocity"], ["velocity"], [2], {},
                       orc.one_hot(nt, 7, 1), ef, s, r, vm, inflow, gt, 0.01, dtype)
    return cfg, p, make, n_norms, N


def test_rhs_vjp_matches_finite_differences():
    cfg, p, make, _, N = _problem()
    rhs = make(p)
    rng = np.random.default_rng(1)
    x, lam = rhs.gt[2] + 0.01 * rng.normal(size=(N, 2)), rng.normal(size=(N, 2))
    g, dx = rhs.vjp(x, np.float32(0.02), lam)
    assert (dx[rhs.inflow] == 0).all() and np.abs(dx[~rhs.inflow]).max() > 0
    for _ in range(6):                               # d/dx
        v = rng.normal(size=x.shape)
        fd = ((rhs(x + 1e-6 * v, 0.02) - rhs(x - 1e-6 * v, 0.02)) * lam).sum() / 2e-6
        assert abs(fd - (dx * v).sum()) < 1e-6 * max(1.0, abs(fd))
    for _ in range(4):                               # d/dparams
        v = rng.normal(size=p.shape)
        fd = ((make(p + 1e-6 * v)(x, 0.02) - make(p - 1e-6 * v)(x, 0.02)) * lam).sum() / 2e-6
        assert abs(fd - (g * v).sum()) < 1e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("solver,n_sub", [("euler", 1), ("euler", 2), ("rk4", 1), ("tsit5", 1)])
def test_multiple_shooting_gradient_matches_finite_differences(solver, n_sub):
    cfg, p, make, _, N = _problem(seed=2)
    args = dict(tstart=0.0, dt=0.01, tstop=0.06, interval_size=3, continuity_term=100, solver=solver, n_sub=n_sub)
    g, loss, preds = sol.train_step_multiple_shooting(make(p), **args)
    assert [q.shape[0] for q in preds] == [3, 3, 3] and loss > 0
    rng = np.random.default_rng(3)
    for _ in range(4):
        v = rng.normal(size=p.shape)
        lp = sol.train_step_multiple_shooting(make(p + 1e-7 * v), **args)[1]
        lm = sol.train_step_multiple_shooting(make(p - 1e-7 * v), **args)[1]
        fd = (lp - lm) / 2e-7
        assert abs(fd - (g * v).sum()) < 2e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("solver", ["euler", "tsit5"])
def test_solver_training_gradient_matches_finite_differences(solver):
    cfg, p, make, n_norms, N = _problem(seed=4, T=5)
    args = dict(n_norms=n_norms, target_fields=["velocity"], target_dims=[2], tstart=0.0, dt=0.01, tstop=0.04,
                solver=solver)
    g, loss, pred = sol.train_step_solver_training(make(p), **args)
    assert pred.shape == (5, N, 2) and np.array_equal(pred[0], make(p).gt[0])
    rng = np.random.default_rng(5)
    for _ in range(4):
        # the random-init rollout is strongly nonlinear in the parameters (loss ~ 1e3 after four Tsit5
        # steps): the central difference only converges for very small perturbations
        v = rng.normal(size=p.shape)
        lp = sol.train_step_solver_training(make(p + 1e-8 * v), **args)[1]
        lm = sol.train_step_solver_training(make(p - 1e-8 * v), **args)[1]
        fd = (lp - lm) / 2e-8
        assert abs(fd - (g * v).sum()) < 2e-5 * max(1.0, abs(fd))


def test_one_interval_shooting_equals_unnormalised_solver_loss():
    """MultipleShooting with a single interval is SolverTraining without the normaliser in the loss."""
    cfg, p, make, n_norms, N = _problem(seed=6, T=5)
    ident = {"velocity": orc.NormaliserOfflineMinMax(0.0, 1.0)}
    g1, l1, _ = sol.train_step_solver_training(make(p), ident, ["velocity"], [2], 0.0, 0.01, 0.04)
    g2, l2, _ = sol.train_step_multiple_shooting(make(p), 0.0, 0.01, 0.04, interval_size=5)
    assert abs(l1 - l2) < 1e-14 * max(1, abs(l1)) and np.allclose(g1, g2, rtol=1e-12, atol=1e-16)
