"""Solver-based training (src/strategies.jl SolverStrategy: SolverTraining :229-286, MultipleShooting :310-386)
over the NeuralODE right-hand side ode_func_train (src/solve.jl:101-115), B200-first:

* **All shooting intervals advance in lock-step as ONE block-diagonal graph** (K copies of the mesh, node ids shifted
  by k*N).  The reference solves the intervals one after the other (strategies.jl:349-362), each a launch-bound
  forward on a 1.9k-node mesh; here one right-hand-side evaluation covers K intervals and runs in the same
  full-occupancy regime as the batched derivative-training step.
* The integrator is a fixed-step explicit Runge-Kutta method (`solargs`: adaptive = false, dt = h - the Euler
  configuration of examples/cylinder_flow/cylinder_flow.jl:79-84; Tsit5 / RK4 tableaus with the same fixed step), and
  the gradient is the exact reverse sweep of that discrete computation: states are checkpointed at every step, the
  stages of a step are recomputed with their activations kept side by side in per-stage workspaces, and every stage
  is pulled back through mgn_backward (d_params and d_nf - what ZygoteVJP asks of ode_step, strategies.jl:183-194).
  The reference's InterpolatingAdjoint(checkpointing = true) is the continuous-time limit of this sweep.
* Normaliser statistics are frozen during the step (K lock-step intervals cannot reproduce K sequential
  accumulate-then-normalise calls; see oracle/mgn_oracle_solver.py).

The engine is written against two small interfaces - `rhs` (forward / backward of the batched right-hand side) and
`alg` (the elementwise / loss kernels) - whose only product implementations are DeviceRhs and DeviceAlgebra below:
every arithmetic operation is a libmgn_b200 kernel; torch allocates, indexes and stacks.  There is no CPU path."""
from __future__ import annotations

import ctypes as C
from fractions import Fraction

import numpy as np
import torch

from . import _lib
from ._lib import MgnError, call
from .core import FeatureGraph, _dev_f32, _ptr, _stream

# (c_2.., rows of A below the diagonal, b).  Tsit5: Tsitouras 2011 (OrdinaryDiffEq.Tsit5, src/solve.jl:58 default).
RK_TABLEAUS = {
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


def time_steps(tstart, dt, tstop):
    """`tstart:dt:tstop` of Float32 fields (src/strategies.jl:344): Julia lifts the endpoints to nearby simple rationals
    so that the range hits `tstop` exactly; element i is Float32(tstart + i*dt)."""
    a, d, b = (Fraction(float(np.float32(v))).limit_denominator(1000000) for v in (tstart, dt, tstop))
    if d <= 0 or b < a:
        raise ValueError("time range needs dt > 0 and tstop >= tstart")
    n = int((b - a) // d) + 1
    return np.asarray([np.float32(float(a + i * d)) for i in range(n)], dtype=np.float32)


def shooting_ranges(n_tsteps, interval_size):
    """src/strategies.jl:346-347, 0-based (first, last inclusive): consecutive intervals share their boundary point."""
    if interval_size < 2:
        raise ValueError("interval_size must be at least 2")
    return [(i - 1, min(n_tsteps, i + interval_size - 1) - 1) for i in range(1, n_tsteps, interval_size - 1)]


def shard_intervals(n_intervals, rank, world):
    """Shooting intervals are independent solves: rank r integrates intervals r, r + world, ...; the
    continuity term between intervals i-1 and i only involves the prediction of i-1 and data, so it belongs to the
    owner of i-1.  Loss and gradient are then SUMMED over ranks."""
    return list(range(rank, n_intervals, world))


def inflow_index(t, strategy_dt, n_data):
    """`floor(Int, t / strategy.dt) + 1` in Float32 (src/solve.jl:106), 0-based, clamped to the data (the reference
    would throw a BoundsError past the end)."""
    return min(int(np.floor(np.float32(t) / np.float32(strategy_dt))), n_data - 1)


# ------------------------------------------------------------------------------------------------------------------
# Device implementations of the two interfaces
# ------------------------------------------------------------------------------------------------------------------

_mesh_cache: dict = {}


def _block_diagonal(senders, receivers, n_nodes, K):
    """Index vectors of K disjoint copies of the mesh (node ids of copy k shifted by k*N, index base kept).  Cached per
    (senders, receivers, K): every training step of a trajectory then reuses the same tensors, hence the same device
    CSR (core.graph_index_for keys on their addresses) and the same workspaces."""
    key = (senders.data_ptr(), receivers.data_ptr(), int(n_nodes), int(senders.shape[0]), int(K))
    hit = _mesh_cache.get(key)
    if hit is None:
        if len(_mesh_cache) >= 4:
            _mesh_cache.pop(next(iter(_mesh_cache)))
        shift = (torch.arange(K, device=senders.device, dtype=torch.int32) * int(n_nodes)).repeat_interleave(
            int(senders.shape[0]))
        # the entry keeps the source tensors alive so that their addresses cannot be reused by another mesh
        hit = ((senders.repeat(K) + shift).contiguous(), (receivers.repeat(K) + shift).contiguous(), senders, receivers)
        _mesh_cache[key] = hit
    return hit[0], hit[1]



class DeviceAlgebra:
    """Elementwise and loss kernels of libmgn_b200 (include/mgn_b200.h, "NeuralODE callers")."""

    def lincomb(self, x, ks, coefs, out=None):
        """out = x + sum_j coefs[j]*ks[j] (x may be None = 0); out may alias x or any ks[j]."""
        terms = [(k, float(c)) for k, c in zip(ks, coefs) if c != 0.0]
        ref = x if x is not None else ks[0]
        if out is None:
            out = torch.empty_like(ref)
        for lo in range(0, max(len(terms), 1), 8):
            part = terms[lo:lo + 8]
            arr_k = (C.c_void_p * max(len(part), 1))(*[k.data_ptr() for k, _ in part])
            arr_c = (C.c_float * max(len(part), 1))(*[c for _, c in part])
            src = x if lo == 0 else out
            call("mgn_ode_lincomb", _ptr(src), arr_k, arr_c, len(part), out.numel(), _ptr(out), _stream())
        return out

    def overwrite(self, x, src, mask):
        out = torch.empty_like(x)
        call("mgn_masked_overwrite", _ptr(x), _ptr(src), _ptr(mask), x.numel(), _ptr(out), _stream())
        return out

    def mul(self, a, b):
        out = torch.empty_like(a)
        call("mgn_vec_mul", _ptr(a), _ptr(b), a.numel(), _ptr(out), _stream())
        return out

    def mse(self, pred, gt, vm, weight, accumulate, loss, dpred):
        """loss (+)= weight * sum((gt - pred)^2 * vm), dpred = d loss / d pred; pred/gt/dpred [n_saves, N, S] views."""
        call("mgn_shooting_mse", _ptr(pred), _ptr(gt), _ptr(vm), pred.shape[0], vm.numel(), float(weight),
             int(bool(accumulate)), _ptr(loss), _ptr(dpred), _stream())

    def continuity(self, a, b, weight, loss, da):
        call("mgn_shooting_continuity", _ptr(a), _ptr(b), a.numel(), float(weight), _ptr(loss), _ptr(da), _stream())


class DeviceRhs:
    """ode_func_train (src/solve.jl:101-115: inflow overwrite, then ode_step :188-219) for K shooting intervals at once,
    and its pullback.  State matrices are [K*N, S]."""

    def __init__(self, mgn, ps, fields, target_fields, target_dict, inputs, node_type, edge_feats, senders, receivers,
                 val_mask, inflow_mask, gt, n_intervals, alg=None):
        self.mgn, self.ps, self.alg = mgn, ps, alg or DeviceAlgebra()
        self.fields, self.tf = list(fields), list(target_fields)
        self.K, self.N = int(n_intervals), int(node_type.shape[0])
        self.cols, off = {}, 0
        for f in self.tf:
            self.cols[f] = (off, int(target_dict[f]))
            off += int(target_dict[f])
        self.S = off
        K, N = self.K, self.N
        self.senders, self.receivers = _block_diagonal(senders, receivers, N, K)
        self.node_type = node_type.repeat(K, 1).contiguous()
        self.inputs = {f: _dev_f32(inputs[f], f).repeat(K, 1).contiguous() for f in self.fields if f not in self.cols}
        self.widths = [self.cols[f][1] if f in self.cols else self.inputs[f].shape[1] for f in self.fields]
        self.val_mask = _dev_f32(val_mask, "val_mask").repeat(K, 1).contiguous()
        self.inflow = None
        if inflow_mask is not None and bool(inflow_mask.any()):
            self.inflow = inflow_mask.to(torch.uint8).repeat(K, 1).contiguous()
        self.gt = gt                                            # [T, N, S]
        # build_graph's edge branch (src/graph.jl:93) is the same for every evaluation under frozen statistics
        ef_all = _dev_f32(edge_feats, "edge_features").repeat(K, 1).contiguous()
        self.ef = mgn.e_norm.apply_ld(ef_all, 0, ef_all.shape[1], _lib.NORM_FORWARD, torch.empty_like(ef_all), 0)
        self._saved = {}
        self.n_evals = 0

    def stage_workspace_bytes(self):
        g = FeatureGraph(torch.empty((self.K * self.N, 1), device=self.ef.device), self.ef, self.senders,
                         self.receivers)
        return self.mgn.model.workspace_bytes(g.index, True)

    def state_norm(self, x, mode, out=None):
        """n_norm[tf] per target field on the columns of a [rows, S] matrix (train_loss(::SolverTraining),
        src/strategies.jl:263-272) or its transposed Jacobian."""
        out = torch.empty_like(x) if out is None else out
        for f, (off, d) in self.cols.items():
            self.mgn.n_norm[f].apply_ld(x, off, d, mode, out, off)
        return out

    def forward(self, x, data_idx, training=False, slot=0):
        """f(x, t): x [K*N, S]; data_idx[k] = 0-based time index of the inflow data of interval k."""
        self.n_evals += 1
        mgn, K, N = self.mgn, self.K, self.N
        xin = x
        if self.inflow is not None:
            idx = torch.as_tensor(np.asarray(data_idx, dtype=np.int64), device=x.device)
            src = self.gt.index_select(0, idx).reshape(K * N, self.S)
            xin = self.alg.overwrite(x, src, self.inflow)
        nt_w = self.node_type.shape[1]
        nf = torch.empty((K * N, sum(self.widths) + nt_w), dtype=torch.float32, device=x.device)
        mgn.n_norm["node_type"].apply_ld(self.node_type, 0, nt_w, _lib.NORM_FORWARD, nf, sum(self.widths))
        col = 0
        for f, w in zip(self.fields, self.widths):
            if f in self.cols:
                mgn.n_norm[f].apply_ld(xin, self.cols[f][0], w, _lib.NORM_FORWARD, nf, col)
            else:
                mgn.n_norm[f].apply_ld(self.inputs[f], 0, w, _lib.NORM_FORWARD, nf, col)
            col += w
        graph = FeatureGraph(nf, self.ef, self.senders, self.receivers)
        out = mgn.model.forward(graph, self.ps, training=training, slot=slot)
        buf = torch.empty_like(out)
        for f, (off, d) in self.cols.items():
            mgn.o_norm[f].apply_ld(out, off, d, _lib.NORM_INVERSE, buf, off)
        if training:
            self._saved[slot] = graph
        return self.alg.mul(buf, self.val_mask)

    def backward(self, dy, slot=0):
        """Pullback of the matching forward(training=True, slot): -> (d_params, d_x)."""
        graph = self._saved.pop(slot, None)
        if graph is None:
            raise MgnError(-1, f"DeviceRhs.backward: no training forward saved in slot {slot}")
        mgn = self.mgn
        dbuf = self.alg.mul(dy, self.val_mask)
        dout = torch.empty_like(dbuf)
        for f, (off, d) in self.cols.items():
            mgn.o_norm[f].apply_ld(dbuf, off, d, _lib.NORM_INVERSE_VJP, dout, off)
        dps, dnf = mgn.model.backward(graph, self.ps, dout, want_dnf=True, slot=slot)
        dxin = torch.zeros_like(dbuf) if any(f not in self.fields for f in self.cols) else torch.empty_like(dbuf)
        col = 0
        for f, w in zip(self.fields, self.widths):
            if f in self.cols:
                mgn.n_norm[f].apply_ld(dnf, col, w, _lib.NORM_FORWARD_VJP, dxin, self.cols[f][0])
            col += w
        if self.inflow is not None:
            return dps, self.alg.overwrite(dxin, None, self.inflow)
        return dps, dxin


# ------------------------------------------------------------------------------------------------------------------
# The engine (generic over rhs / alg)
# ------------------------------------------------------------------------------------------------------------------


class ShootingEngine:
    """Lock-step fixed-step solve of K intervals and its reverse sweep."""

    def __init__(self, rhs, alg, solver="euler", n_sub=1, strategy_dt=0.01, n_data=None, stage_slots=True):
        if solver not in RK_TABLEAUS:
            raise MgnError(-1, f"unknown fixed-step solver {solver!r}: one of {sorted(RK_TABLEAUS)}")
        self.rhs, self.alg = rhs, alg
        self.c, self.A, self.b = RK_TABLEAUS[solver]
        self.n_sub, self.sdt, self.n_data = int(n_sub), np.float32(strategy_dt), n_data
        self.stage_slots = bool(stage_slots)

    def _idx(self, tvec):
        return [inflow_index(t, self.sdt, self.n_data) for t in tvec]

    def _stages(self, x, tvec, h, training):
        """Stage inputs, stage time vectors and slopes of one step (slot i holds stage i when training)."""
        xs, ts = [x], [tvec]
        ks = [self.rhs.forward(x, self._idx(tvec), training=training, slot=0)]
        for i, (ci, a) in enumerate(zip(self.c, self.A), start=1):
            xi = self.alg.lincomb(x, ks, [np.float32(h) * np.float32(aj) for aj in a])
            ti = [np.float32(t + np.float32(ci) * h) for t in tvec]
            xs.append(xi)
            ts.append(ti)
            ks.append(self.rhs.forward(xi, self._idx(ti), training=training, slot=i if training else 0))
        return xs, ts, ks

    def solve(self, x0, macro_times, dt):
        """x0 [K*N, S]; macro_times[m][k] = Float32 time of interval k at its m-th save.  Returns the saved states
        (len(macro_times) of them, x0 first), the per-step checkpoints and h."""
        h = np.float32(np.float32(dt) / np.float32(self.n_sub))
        x, saves, chk = x0, [x0], []
        for m in range(len(macro_times) - 1):
            for j in range(self.n_sub):
                tvec = [np.float32(t + np.float32(j) * h) for t in macro_times[m]]
                chk.append((x, tvec))
                _, _, ks = self._stages(x, tvec, h, training=False)
                x = self.alg.lincomb(x, ks, [np.float32(h) * np.float32(bi) for bi in self.b])
            saves.append(x)
        return saves, chk, h

    def adjoint(self, chk, h, dsaves, n_params_like):
        """Reverse sweep: dsaves[m] = d loss / d saves[m] ([K*N, S]).  Returns d loss / d params (the gradient w.r.t.
        the initial states is dropped: they are data)."""
        s = len(self.b)
        g = torch.zeros_like(n_params_like)
        lam = dsaves[-1].clone()
        for n in range(len(chk) - 1, -1, -1):
            x, tvec = chk[n]
            slots = self.stage_slots and s > 1
            xs, ts, _ = self._stages(x, tvec, h, training=slots) if s > 1 else ([x], [tvec], None)
            dk = [self.alg.lincomb(None, [lam], [np.float32(h) * np.float32(bi)]) for bi in self.b]
            for i in range(s - 1, -1, -1):
                if not slots:
                    self.rhs.forward(xs[i], self._idx(ts[i]), training=True, slot=0)
                gi, dxi = self.rhs.backward(dk[i], slot=i if slots else 0)
                self.alg.lincomb(g, [gi], [1.0], out=g)
                self.alg.lincomb(lam, [dxi], [1.0], out=lam)
                if i > 0:
                    for j, aij in enumerate(self.A[i - 1]):
                        if aij != 0.0:
                            self.alg.lincomb(dk[j], [dxi], [np.float32(h) * np.float32(aij)], out=dk[j])
            if n % self.n_sub == 0 and n > 0:
                self.alg.lincomb(lam, [dsaves[n // self.n_sub]], [1.0], out=lam)
        return g


def multiple_shooting_step(rhs, alg, ps, gt, val_mask, tstart, dt, tstop, interval_size, continuity_term=100,
                           solver="euler", n_sub=1, owned=None, stage_slots=True):
    """train_step + train_loss(::MultipleShooting) (src/strategies.jl:174-199, :343-386) -> (gs, loss, preds).
    `rhs` must have been built for len(owned) intervals (all of them when owned is None).  gt [T, N, S]."""
    ts = time_steps(tstart, dt, tstop)
    ranges = shooting_ranges(len(ts), interval_size)
    owned = list(range(len(ranges))) if owned is None else list(owned)
    K, N, S = len(owned), gt.shape[1], gt.shape[2]
    if K == 0:
        return torch.zeros_like(ps), torch.zeros(1, dtype=gt.dtype, device=ps.device), []
    if ranges[-1][1] >= gt.shape[0]:
        raise ValueError(f"tstart:dt:tstop has {len(ts)} points but the trajectory only {gt.shape[0]}")
    M = max(ranges[i][1] - ranges[i][0] for i in owned)
    firsts = [ranges[i][0] for i in owned]
    lens = [ranges[i][1] - ranges[i][0] + 1 for i in owned]
    # interval k at macro step m sits at tsteps[first_k + m]; intervals shorter than M keep stepping past their end
    # (their extra states enter no loss term, so they receive a zero cotangent)
    macro = [[ts[min(f + m, len(ts) - 1)] for f in firsts] for m in range(M + 1)]
    eng = ShootingEngine(rhs, alg, solver, n_sub, strategy_dt=dt, n_data=gt.shape[0], stage_slots=stage_slots)
    dev = gt.device
    x0 = gt.index_select(0, torch.as_tensor(firsts, device=dev)).reshape(K * N, S).contiguous()
    saves, chk, h = eng.solve(x0, macro, dt)
    P = torch.stack(saves).reshape(M + 1, K, N, S).permute(1, 0, 2, 3).contiguous()          # [K, M+1, N, S]
    gidx = torch.as_tensor([[min(f + m, gt.shape[0] - 1) for m in range(M + 1)] for f in firsts], device=dev)
    G = gt.index_select(0, gidx.reshape(-1)).reshape(K, M + 1, N, S).contiguous()
    dP = torch.zeros_like(P)
    loss = torch.zeros(1, dtype=gt.dtype, device=dev)
    for k, (i, ln) in enumerate(zip(owned, lens)):
        alg.mse(P[k, :ln], G[k, :ln], val_mask, 1.0 / (ln * N * S), True, loss, dP[k, :ln])
        if i + 1 < len(ranges):                                  # continuity term of interval i+1, owned by i's owner
            alg.continuity(P[k, ln - 1], gt[ranges[i + 1][0]], float(continuity_term), loss, dP[k, ln - 1])
    dsaves = list(dP.permute(1, 0, 2, 3).contiguous().reshape(M + 1, K * N, S))
    g = eng.adjoint(chk, h, dsaves, ps)
    return g, loss, [P[k, :ln] for k, ln in enumerate(lens)]


def solver_training_step(rhs, alg, ps, gt, val_mask, tstart, dt, tstop, solver="euler", n_sub=1, stage_slots=True):
    """train_step + train_loss(::SolverTraining) (src/strategies.jl:174-199, :253-286) -> (gs, loss, pred [T', N, S])."""
    ts = time_steps(tstart, dt, tstop)
    Tn, N, S = len(ts), gt.shape[1], gt.shape[2]
    if Tn > gt.shape[0]:
        raise ValueError(f"tstart:dt:tstop has {Tn} points but the trajectory only {gt.shape[0]}")
    eng = ShootingEngine(rhs, alg, solver, n_sub, strategy_dt=dt, n_data=gt.shape[0], stage_slots=stage_slots)
    saves, chk, h = eng.solve(gt[0].contiguous(), [[t] for t in ts], dt)
    pred = torch.stack(saves)                                                                # [T', N, S]
    flat = lambda a: a.reshape(Tn * N, S)
    pred_n = rhs.state_norm(flat(pred), _lib.NORM_FORWARD)
    gt_n = rhs.state_norm(flat(gt[:Tn].contiguous()), _lib.NORM_FORWARD)
    loss = torch.zeros(1, dtype=gt.dtype, device=gt.device)
    dpred_n = torch.empty_like(pred_n)
    alg.mse(pred_n.reshape(Tn, N, S), gt_n.reshape(Tn, N, S), val_mask, 1.0 / (Tn * N * S), False, loss,
            dpred_n.reshape(Tn, N, S))
    dpred = rhs.state_norm(dpred_n, _lib.NORM_FORWARD_VJP).reshape(Tn, N, S)
    g = eng.adjoint(chk, h, list(dpred), ps)
    return g, loss, pred
