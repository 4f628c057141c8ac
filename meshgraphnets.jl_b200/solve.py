"""Mirror of src/solve.jl: ode_step (:188-219), ode_func_eval (:147-158) and rollout (:42-68) with the
fixed-step Euler configuration of examples/cylinder_flow/cylinder_flow.jl:79-84.  The adaptive
Tsit5 driver itself is OrdinaryDiffEq's (out of scope); the fixed-step tableaus of shooting.RK_TABLEAUS restate
one explicit step so that the 6-RHS-evaluations-per-step workload of config 3 can be run and timed.

Every arithmetic operation - inflow overwrite, normalisers, model, inverse_data, val_mask, the Runge-Kutta
combinations - is a libmgn_b200 kernel that only enqueues on the current stream, so a whole rollout can be captured
once and replayed as ONE CUDA graph (CapturedRollout): no host round trip per stage."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MgnError
from .fused import FusedGraph, forward_fused
from .graph import build_graph
from .shooting import RK_TABLEAUS, DeviceAlgebra

_alg = DeviceAlgebra()


def ode_step(x, p, t):
    """src/solve.jl:188-219.  x is [N, sum(target dims)]; p = (mgn, ps, inputs, fields, meta,
    target_fields, target_dict, node_type, edge_features, senders, receivers, val_mask).
    ONE fused model call (mgn_forward_fused): the state columns, the other input fields and the one-hot node types are
    normalised and concatenated while the encoder stages its operand, `inverse_data(o_norm, out) .* val_mask` runs in
    the decoder's epilogue; the normalisers' accumulate branch (they still fire in the reference's rollout) is two
    launches for all of them.  ode_step_unfused is the operation-by-operation form (same values)."""
    mgn, ps, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders, receivers, val_mask = p
    x = x.contiguous()
    offs, off = {}, 0
    for k in target_fields:
        offs[k] = off
        off += target_dict[k]
    blocks = []
    for f in fields:
        if f in offs:
            blocks.append((mgn.n_norm[f], x, offs[f], target_dict[f]))
        else:
            v = inputs[f]
            v = (v[0] if v.dim() == 3 else v).contiguous()
            blocks.append((mgn.n_norm[f], v, 0, v.shape[1]))
    # the reference evaluates the node_type normaliser first (graph.jl:80); it lands last (graph.jl:86)
    blocks.append((mgn.n_norm["node_type"], node_type, 0, node_type.shape[1]))
    out_blocks = [(mgn.o_norm[tf], int(meta["features"][tf]["dim"])) for tf in target_fields]
    fg = FusedGraph(blocks, (mgn.e_norm, edge_feats, 0, edge_feats.shape[1]), senders, receivers, node_type.shape[0],
                    out_blocks, val_mask)
    fg.accumulate()
    return forward_fused(mgn.model, fg, ps)


def ode_step_unfused(x, p, t):
    """ode_step as the sequence of separate operations of src/solve.jl:188-219 (build_graph, model, inverse_data per
    target field, val_mask product)."""
    mgn, ps, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders, receivers, val_mask = p
    offset = 0
    for k in target_fields:
        inputs[k] = x[:, offset:offset + target_dict[k]].contiguous()
        offset += target_dict[k]
    graph = build_graph(mgn, inputs, fields, 1, node_type, edge_feats, senders, receivers)
    output, st = mgn.model(graph, ps, mgn.st)
    mgn.st = st
    buf = torch.empty_like(output)
    col = 0
    for tf in target_fields:
        d = meta["features"][tf]["dim"]
        mgn.o_norm[tf].inverse(output[:, col:col + d].contiguous(), out=buf, col=col)
        col += d
    return _alg.mul(buf, val_mask)


def _inflow_u8(inflow_mask):
    return inflow_mask if inflow_mask.dtype == torch.uint8 else inflow_mask.to(torch.uint8)


def ode_func_eval(x, p, t):
    """src/solve.jl:147-158: overwrite the inflow nodes from data, then ode_step."""
    (mgn, ps, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
     receivers, val_mask, inflow_mask, saves_dt) = p
    if inflow_mask is not None:
        # floor(Int, t / saves_dt) + 1 with Float32 t and saves_dt (0-based here)
        idx = int(np.floor(np.float32(t) / np.float32(saves_dt)))
        cur = torch.cat([data[f][idx] for f in target_fields], dim=1) if len(target_fields) > 1 \
            else data[target_fields[0]][idx]
        # in place: the reference mutates the solver state
        x.copy_(_alg.overwrite(x, cur.contiguous(), _inflow_u8(inflow_mask)))
    return ode_step(x, (mgn, ps, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats,
                        senders, receivers, val_mask), t)


def rk_step(f, x, t, dt, solver="euler"):
    """One fixed step of an explicit Runge-Kutta method (shooting.RK_TABLEAUS) through mgn_ode_lincomb."""
    c, A, b = RK_TABLEAUS[solver]
    h = np.float32(dt)
    ks = [f(x, t)]
    for ci, a in zip(c, A):
        xi = _alg.lincomb(x, ks, [h * np.float32(aj) for aj in a])
        ks.append(f(xi, np.float32(np.float32(t) + np.float32(ci) * h)))
    return _alg.lincomb(x, ks, [h * np.float32(bi) for bi in b])


def euler_step(f, x, t, dt):
    return rk_step(f, x, t, dt, "euler")


def tsit5_step(f, x, t, dt):
    """One explicit Tsitouras 5(4) step (6 RHS evaluations; the 7th FSAL stage is the next step's
    first) - the per-step RHS workload of OrdinaryDiffEq.Tsit5 at src/solve.jl:58."""
    return rk_step(f, x, t, dt, "tsit5")


def _rollout_params(mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
                    receivers, val_mask, inflow_mask, data, saves):
    inputs = {k: v for k, v in initial_state.items() if k not in target_dict}
    if inflow_mask is not None:
        inflow_mask = _inflow_u8(inflow_mask).contiguous()
    return (mgn, mgn.ps, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
            receivers, val_mask, inflow_mask, saves[1] - saves[0])


def _step_plan(start, stop, dt, saves):
    """The fixed-step grid of ``solve(prob, solver; adaptive=false, dt=dt, saveat=saves)`` over (start, stop)
    (src/solve.jl:62): the integrator advances in steps of ``dt`` from ``start`` and a save time must coincide with a
    grid point (interpolated saves are OrdinaryDiffEq's dense output: out of scope, so they raise instead of silently
    landing on the wrong physical time).  ``dt = None`` is the ``tstops = saves`` branch (:60): one step per save
    interval.  Returns [(n_sub, h, t0)] per save interval (t0 = the Float32 time the interval starts at: sub-step j
    runs at Float32(t0 + j*h), never at an accumulated sum), preceded by the (start -> saves[0]) lead-in."""
    saves = [float(np.float32(v)) for v in saves]
    if len(saves) < 2:
        raise MgnError(-1, "rollout needs at least two save times (saves[2] - saves[1] is the inflow data spacing)")
    start = saves[0] if start is None else float(np.float32(start))
    if start > saves[0] + 1e-6 * max(1.0, abs(saves[0])) or (stop is not None and saves[-1] > float(stop) * (1 + 1e-6) + 1e-12):
        raise MgnError(-1, f"save times {saves[0]}..{saves[-1]} are outside the integration interval ({start}, {stop})")
    plan = []
    for a, b in zip([start] + saves[:-1], saves):
        span = b - a
        if span < -1e-9:
            raise MgnError(-1, "save times must be increasing")
        if dt is None:
            plan.append((1 if span > 1e-9 * max(1.0, abs(b)) else 0, np.float32(span), np.float32(a)))
            continue
        n = int(round(span / float(dt)))
        if abs(n * float(dt) - span) > 1e-4 * float(dt) + 1e-7 * max(1.0, abs(b)):
            raise MgnError(-1, f"the fixed step dt = {dt} does not divide the save interval ({a}, {b}): saves between "
                               "grid points need dense output, which this mirror does not provide")
        plan.append((n, np.float32(dt), np.float32(a)))
    return start, saves, plan


def _integrate(f, x, start, plan, solver, on_save):
    for k, (n_sub, h, t0) in enumerate(plan):
        for j in range(n_sub):
            x = rk_step(f, x, np.float32(t0 + np.float32(j) * h), h, solver)
        on_save(k, x)
    return x


def rollout(mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
            receivers, val_mask, inflow_mask, data, start, stop, dt, saves, solver="euler"):
    """src/solve.jl:42-68 with ``solve(prob, solver; adaptive=false, dt=dt, saveat=saves)`` over (start, stop): fixed
    steps of ``dt`` (several per save interval when dt divides it), inflow data indexed by the sub-step's own time.
    Returns (list of saved states, times)."""
    x = torch.cat([initial_state[f] for f in target_fields], dim=1).clone()
    p = _rollout_params(mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
                        receivers, val_mask, inflow_mask, data, saves)
    start, saves_f, plan = _step_plan(start, stop, dt, saves)
    sol = []
    _integrate(lambda xx, tt: ode_func_eval(xx, p, tt), x, start, plan, solver, lambda k, xx: sol.append(xx.clone()))
    return sol, saves_f


class CapturedRollout:
    """The same rollout captured once as a single CUDA graph and replayed: ``replay(initial_state)`` copies the
    initial state into the graph's input and returns the saved states (tensors owned by the graph, overwritten by the
    next replay).  Normaliser statistics must not change between capture and replay (they are read on the device,
    so accumulated statistics ARE seen; only the accumulate branch itself is skipped once max_acc is reached)."""

    def __init__(self, mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
                 receivers, val_mask, inflow_mask, data, start, stop, dt, saves, solver="euler"):
        self.target_fields = list(target_fields)
        self.x0 = torch.cat([initial_state[f] for f in target_fields], dim=1).clone()
        args = (mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders, receivers,
                val_mask, inflow_mask, data, saves)
        p = _rollout_params(*args)
        self._keep = p      # the graph holds raw addresses: every tensor it reads must outlive it (e.g. the uint8 mask)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        start, saves_f, plan = _step_plan(start, stop, dt, saves)
        h0 = next((h for n, h, _ in plan if n > 0), np.float32(saves[1] - saves[0]))
        with torch.cuda.stream(side):             # warm-up: allocations and first-use scratch happen outside capture
            for _ in range(2):
                rk_step(lambda xx, tt: ode_func_eval(xx, p, tt), self.x0.clone(), np.float32(start), h0, solver)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.sol = []
            _integrate(lambda xx, tt: ode_func_eval(xx, p, tt), self.x0.clone(), start, plan, solver,
                       lambda k, xx: self.sol.append(xx.clone()))
        self.ts = saves_f

    def replay(self, initial_state=None):
        if initial_state is not None:
            self.x0.copy_(torch.cat([initial_state[f] for f in self.target_fields], dim=1))
        self.graph.replay()
        return self.sol, self.ts
