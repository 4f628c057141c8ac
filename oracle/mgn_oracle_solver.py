"""CPU oracle for the solver-based training strategies of src/strategies.jl (SolverTraining,
MultipleShooting) over the NeuralODE right-hand side of src/solve.jl.

TEST INFRASTRUCTURE ONLY (same rule as mgn_oracle.py): imported by tests/ only.

PARITY UNPINNED, and one step further from the reference than mgn_oracle.py: the reference hands
the right-hand side to OrdinaryDiffEq (`solve(..., strategy.solver; sensealg = InterpolatingAdjoint(
autojacvec = ZygoteVJP(), checkpointing = true))`, src/strategies.jl:253-259, :349-362) - an adaptive
integrator and a continuous adjoint whose results depend on tolerances.  What is restated here is the
configuration that has a closed form: a FIXED-STEP explicit Runge-Kutta method (`adaptive = false,
dt = h` in `solargs`, as examples/cylinder_flow/cylinder_flow.jl:79-84 does for Euler) and the exact
gradient of that discrete computation (discretise-then-optimise), which the continuous adjoint
approaches as h -> 0.  The gradient code below is checked against finite differences in fp64
(tests/test_oracle_solver.py).

Normaliser statistics are FROZEN during a solver training step (no accumulation inside the
right-hand side): the product evaluates all shooting intervals as one block-diagonal graph, which
cannot reproduce K sequential accumulate-then-normalise calls; after `max_norm_steps`
(src/MeshGraphNets.jl:44) the reference does not accumulate either.

Layout: state x is [N, S] (C view of Julia's S x N), gt is [T, N, S].
"""
from __future__ import annotations

import numpy as np

import mgn_oracle as orc

# Butcher tableaus (c_2.., rows of A below the diagonal, b).  Tsit5: Tsitouras 2011, the method behind
# OrdinaryDiffEq.Tsit5 (src/solve.jl:58 default); the 7th FSAL stage only feeds the error estimate.
TABLEAUS = {
    "euler": ((), (), (1.0,)),
    "rk4": ((0.5, 0.5, 1.0), ((0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)), (1 / 6, 1 / 3, 1 / 3, 1 / 6)),
    "tsit5": (
        (0.161, 0.327, 0.9, 0.9800255409045097, 1.0),
        ((0.161,),
         (-0.008480655492356989, 0.335480655492357),
         (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
         (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
         (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383)),
        (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)),
}


def shooting_ranges(n_tsteps, interval_size):
    """src/strategies.jl:346-347: `[i:min(length(tsteps), i + interval_size - 1) for i in
    1:(interval_size - 1):(length(tsteps) - 1)]`, returned 0-based as (first, last_inclusive)."""
    out = []
    for i in range(1, n_tsteps, interval_size - 1):
        out.append((i - 1, min(n_tsteps, i + interval_size - 1) - 1))
    return out


def tsteps(tstart, dt, tstop):
    """`tstart:dt:tstop` with Float32 endpoints (struct fields of src/strategies.jl:229-236).  Julia
    lifts float range arguments to nearby simple rationals (Base.rat / twice precision) so that
    `(0f0:0.01f0:0.49f0)[50] == 0.49f0`; restated with limit_denominator on the Float32 values."""
    from fractions import Fraction
    a, d, b = (Fraction(float(np.float32(v))).limit_denominator(1000000) for v in (tstart, dt, tstop))
    n = int((b - a) // d) + 1
    return np.asarray([np.float32(float(a + i * d)) for i in range(n)], dtype=np.float32)


def affine_of(norm):
    """(a, c) with norm(x) == a x + c for the current (frozen) statistics, fp64."""
    if isinstance(norm, orc.NormaliserOnline):
        sd = norm.std().astype(np.float64)
        return 1.0 / sd, -norm.mean().astype(np.float64) / sd
    if isinstance(norm, orc.NormaliserOfflineMeanStd):
        return 1.0 / np.float64(norm.std), -np.float64(norm.mean) / np.float64(norm.std)
    a = (np.float64(norm.target_max) - np.float64(norm.target_min)) / (np.float64(norm.data_max) - np.float64(norm.data_min))
    return a, np.float64(norm.target_min) - np.float64(norm.data_min) * a


class Rhs:
    """ode_func_train (src/solve.jl:101-115) = inflow overwrite + ode_step (:188-219) with frozen
    normalisers, and its pullback (what ZygoteVJP derives, src/strategies.jl:183-194)."""

    def __init__(self, cfg, params, n_norms, e_norm, o_norms, fields, target_fields, target_dims, inputs,
                 node_type_onehot, edge_feats, senders, receivers, vmask, inflow_mask, gt, strategy_dt,
                 dtype=np.float64):
        self.cfg, self.dtype = cfg, dtype
        self.p = np.asarray(params, dtype)
        self.fields, self.tf, self.td = list(fields), list(target_fields), list(target_dims)
        self.inputs = {k: np.asarray(v, dtype) for k, v in inputs.items()}
        self.s, self.r = senders, receivers
        self.vm = np.asarray(vmask, dtype)
        self.inflow = None if inflow_mask is None else np.asarray(inflow_mask, bool)
        self.gt = np.asarray(gt, dtype)
        self.sdt = strategy_dt
        self.nn = {k: affine_of(v) for k, v in n_norms.items()}
        self.on = {k: affine_of(v) for k, v in o_norms.items()}
        a, c = affine_of(e_norm)
        self.ef = (np.asarray(edge_feats, dtype) * a + c).astype(dtype)
        a, c = self.nn["node_type"]
        self.nt = (np.asarray(node_type_onehot, dtype) * a + c).astype(dtype)
        self.n_evals = 0

    def _cols(self):
        off, out = 0, {}
        for k, d in zip(self.tf, self.td):
            out[k] = (off, d)
            off += d
        return out

    def __call__(self, x, t, tape=None, idx=None):
        """``idx`` overrides the inflow data index derived from t (used by the host-logic test doubles)."""
        self.n_evals += 1
        xin = np.array(x, dtype=self.dtype, copy=True)
        if self.inflow is not None:
            idx = min(orc.inflow_index(t, self.sdt), self.gt.shape[0] - 1) if idx is None else idx
            xin[self.inflow] = self.gt[idx][self.inflow]
        cols = self._cols()
        parts = []
        for f in self.fields:
            v = xin[:, cols[f][0]:cols[f][0] + cols[f][1]] if f in cols else self.inputs[f]
            a, c = self.nn[f]
            parts.append(v * a + c)
        nf = np.concatenate(parts + [self.nt], axis=1).astype(self.dtype)
        out = orc.model_forward(self.cfg, self.p, nf, self.ef, self.s, self.r, 1, tape, self.dtype)
        buf = np.empty_like(out)
        for k, (off, d) in cols.items():
            a, c = self.on[k]
            buf[:, off:off + d] = (out[:, off:off + d] - c) / a
        return (buf * self.vm).astype(self.dtype)

    def vjp(self, x, t, lam, idx=None):
        """-> (d_params, d_x) of lam . f(x, t)."""
        tape = []
        self(x, t, tape, idx)
        cols = self._cols()
        dbuf = np.asarray(lam, self.dtype) * self.vm
        dout = np.empty_like(dbuf)
        for k, (off, d) in cols.items():
            dout[:, off:off + d] = dbuf[:, off:off + d] / self.on[k][0]
        g, dnf = orc.model_backward(self.cfg, self.p, tape, dout, self.s, self.r, x.shape[0], 1)
        dx = np.zeros_like(dbuf)
        col = 0
        for f in self.fields:
            w = cols[f][1] if f in cols else self.inputs[f].shape[1]
            if f in cols:
                dx[:, cols[f][0]:cols[f][0] + w] = dnf[:, col:col + w] * self.nn[f][0]
            col += w
        if self.inflow is not None:
            dx[self.inflow] = 0
        return g, dx


def rk_stage_inputs(f, x, t, h, tab):
    """Stage inputs, stage times and slopes of one explicit Runge-Kutta step."""
    c, A, b = tab
    dt_ = x.dtype.type
    xs, ts, ks = [x], [np.float32(t)], [f(x, np.float32(t))]
    for ci, a in zip(c, A):
        xi = x
        for aj, kj in zip(a, ks):
            if aj != 0.0:
                xi = xi + dt_(np.float32(h) * np.float32(aj)) * kj
        ti = np.float32(np.float32(t) + np.float32(ci) * np.float32(h))
        xs.append(xi)
        ts.append(ti)
        ks.append(f(xi, ti))
    return xs, ts, ks


def rk_step(f, x, t, h, tab):
    xs, ts, ks = rk_stage_inputs(f, x, t, h, tab)
    out = x
    for bj, kj in zip(tab[2], ks):
        if bj != 0.0:
            out = out + x.dtype.type(np.float32(h) * np.float32(bj)) * kj
    return out


def solve_fixed(f, x0, t0, n_saves, dt, n_sub, tab):
    """Saved states at t0, t0 + dt, ... (n_saves of them) with n_sub fixed steps of h = dt/n_sub between
    saves; also returns every sub-step state (the checkpoints of the reverse sweep)."""
    h = np.float32(np.float32(dt) / np.float32(n_sub))
    x = np.array(x0, copy=True)
    saves, chk = [x], []
    for m in range(n_saves - 1):
        tm = np.float32(np.float32(t0) + np.float32(m) * np.float32(dt)) if not isinstance(t0, np.ndarray) else t0[m]
        for j in range(n_sub):
            t = np.float32(tm + np.float32(j) * h)
            chk.append((x, t))
            x = rk_step(f, x, t, h, tab)
        saves.append(x)
    return saves, chk, h


def solve_adjoint(rhs: Rhs, chk, h, n_sub, tab, dsaves):
    """Reverse sweep of solve_fixed: dsaves[m] = d loss / d saves[m]; returns d_params (the gradient
    w.r.t. the initial state is dropped: u0 is data)."""
    c, A, b = tab
    s = len(b)
    g = np.zeros_like(rhs.p)
    lam = np.array(dsaves[-1], copy=True)
    dt_ = lam.dtype.type
    for n in range(len(chk) - 1, -1, -1):
        x, t = chk[n]
        xs, ts, _ = rk_stage_inputs(rhs, x, t, h, tab)
        dk = [dt_(np.float32(h) * np.float32(bi)) * lam for bi in b]
        for i in range(s - 1, -1, -1):
            gi, dxi = rhs.vjp(xs[i], ts[i], dk[i])
            g += gi
            lam = lam + dxi
            if i > 0:
                for j, aij in enumerate(A[i - 1]):
                    if aij != 0.0:
                        dk[j] = dk[j] + dt_(np.float32(h) * np.float32(aij)) * dxi
        if n % n_sub == 0 and n > 0:
            lam = lam + dsaves[n // n_sub]
    return g


def train_step_solver_training(rhs: Rhs, n_norms, target_fields, target_dims, tstart, dt, tstop, solver="euler",
                               n_sub=1):
    """train_step + train_loss(::SolverTraining), src/strategies.jl:174-199, :253-286 -> (gs, loss)."""
    ts = tsteps(tstart, dt, tstop)
    tab = TABLEAUS[solver]
    saves, chk, h = solve_fixed(rhs, rhs.gt[0], ts, len(ts), dt, n_sub, tab)
    pred = np.stack(saves)                                    # [T', N, S]
    gt = rhs.gt[:pred.shape[0]]
    scale = np.concatenate([np.broadcast_to(affine_of(n_norms[f])[0], (d,)) for f, d in zip(target_fields, target_dims)])
    err = ((gt - pred) * scale) ** 2 * rhs.vm                 # n_norm(gt) - n_norm(pred): the shift cancels
    loss = err.mean()
    dpred = -2.0 * scale * scale * (gt - pred) * rhs.vm / err.size
    return solve_adjoint(rhs, chk, h, n_sub, tab, list(dpred)), loss, pred


def train_step_multiple_shooting(rhs: Rhs, tstart, dt, tstop, interval_size, continuity_term=100, solver="euler",
                                 n_sub=1):
    """train_step + train_loss(::MultipleShooting), src/strategies.jl:174-199, :343-386 -> (gs, loss, preds)."""
    ts = tsteps(tstart, dt, tstop)
    tab = TABLEAUS[solver]
    ranges = shooting_ranges(len(ts), interval_size)
    sols = []
    for (a, b) in ranges:
        saves, chk, h = solve_fixed(rhs, rhs.gt[a], ts[a:b + 1], b - a + 1, dt, n_sub, tab)
        sols.append((np.stack(saves), chk, h))
    loss = 0.0
    dpreds = []
    for i, (a, b) in enumerate(ranges):
        pred = sols[i][0]
        gt = rhs.gt[a:b + 1]
        err = (gt - pred) ** 2 * rhs.vm
        loss += err.mean()
        dpreds.append(-2.0 * (gt - pred) * rhs.vm / err.size)
        if i > 0:
            d = sols[i - 1][0][-1] - rhs.gt[a]
            loss += continuity_term * np.abs(d).sum()
            dpreds[i - 1][-1] = dpreds[i - 1][-1] + continuity_term * np.sign(d)
    g = np.zeros_like(rhs.p)
    for i in range(len(ranges)):
        g += solve_adjoint(rhs, sols[i][1], sols[i][2], n_sub, tab, list(dpreds[i]))
    return g, loss, [s[0] for s in sols]
