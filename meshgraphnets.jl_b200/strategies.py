"""Mirror of the training-strategy entry points of src/strategies.jl that reach the hot path:
DerivativeTraining (:389-447) and the solver strategies SolverTraining (:229-286) and MultipleShooting (:310-386).
`init_train_step` / `train_step` / `get_delta` dispatch on the strategy type as the Julia methods do."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import MgnError
from .core import mse_reduce, step_
from .graph import build_graph
from .shooting import (DeviceAlgebra, DeviceRhs, RK_TABLEAUS, multiple_shooting_step, shard_intervals,
                       shooting_ranges, solver_training_step, time_steps)


class DerivativeTraining:
    """src/strategies.jl:389-447."""

    def __init__(self, window_size=0, random=True):
        self.window_size, self.random = window_size, random


class SolverStrategy:
    """Common part of SolverTraining / MultipleShooting (src/strategies.jl:140, :229-251, :310-341).  `solver` names a
    fixed-step explicit Runge-Kutta method ("euler", "rk4", "tsit5"); `solargs` follow OrdinaryDiffEq's keywords:
    `adaptive` must be false (adaptive stepping is OrdinaryDiffEq's, out of scope) and `dt` is the fixed step h, which
    must divide the save interval.  `sense` is accepted for signature compatibility: the gradient is always the exact
    reverse sweep of the discrete solve (the continuous InterpolatingAdjoint is its h -> 0 limit)."""

    def __init__(self, tstart, dt, tstop, solver, /, sense=None, **solargs):   # positional-only: `dt=h` is a solarg
        self.tstart, self.dt, self.tstop = np.float32(tstart), np.float32(dt), np.float32(tstop)
        self.solver = str(solver).lower()
        if self.solver not in RK_TABLEAUS:
            raise MgnError(-1, f"unknown fixed-step solver {solver!r}: one of {sorted(RK_TABLEAUS)}")
        if solargs.get("adaptive", False):
            raise MgnError(-1, "adaptive step-size control is not provided: pass adaptive=False and dt=h")
        h = np.float32(solargs.get("dt", self.dt))
        self.n_sub = int(round(float(self.dt) / float(h)))
        if self.n_sub < 1 or abs(self.n_sub * float(h) - float(self.dt)) > 1e-6 * float(self.dt):
            raise MgnError(-1, f"the fixed step {h} must divide the save interval {self.dt}")
        self.sense, self.solargs = sense, solargs


class SolverTraining(SolverStrategy):
    """SolverTraining(tstart, dt, tstop, solver; sense, solargs...)  <- src/strategies.jl:229-251."""


class MultipleShooting(SolverStrategy):
    """MultipleShooting(tstart, dt, tstop, solver; sense, interval_size, continuity_term = 100, solargs...)
    <- src/strategies.jl:310-341.  `rank` / `world` shard the intervals over processes (sum loss and gradient)."""

    def __init__(self, tstart, dt, tstop, solver, /, *, interval_size, continuity_term=100, sense=None, rank=0,
                 world=1, **solargs):
        super().__init__(tstart, dt, tstop, solver, sense=sense, **solargs)
        self.interval_size, self.continuity_term = int(interval_size), continuity_term
        self.rank, self.world = int(rank), int(world)
        if self.interval_size < 2:
            raise MgnError(-1, "interval_size must be at least 2 (number of observations in one interval)")


def get_delta(strategy, trajectory_length):
    """src/strategies.jl:142-144 (solver strategies: 1) and :391-393."""
    if isinstance(strategy, SolverStrategy):
        return 1
    return strategy.window_size if strategy.window_size > 0 else trajectory_length - 1


def _init_solver_step(t):
    """src/strategies.jl:146-172: the initial state, the inputs without the targets, gt = vcat(target features) and
    u0 = gt[:, :, 1].  Tensors are [T, N, d] (the C view of Julia's (d, N, T))."""
    mgn, data, meta, fields, target_fields, node_type, edge_feats, senders, receivers, _, _, val_mask = t
    target_dict = {tf: int(meta["features"][tf]["dim"]) for tf in meta["target_features"]}
    initial_state = {k: (v[0] if isinstance(v, torch.Tensor) and v.dim() == 3 else v)
                     for k, v in data.items() if not k.endswith(".ev")}
    inputs = {k: v for k, v in initial_state.items() if not (k.startswith("target|") and k[7:] in target_dict)}
    gts = [data[tf] for tf in meta["target_features"]]
    gt = (torch.cat(gts, dim=2) if len(gts) > 1 else gts[0]).contiguous()
    return (mgn, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders, receivers,
            val_mask, gt[0], gt)


def _solver_train_step(strategy, t):
    """src/strategies.jl:174-199 + train_loss (:253-286 / :343-386)."""
    (mgn, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders, receivers, val_mask,
     u0, gt) = t
    S = sum(target_dict[f] for f in target_fields)
    nt = data["node_type"]
    nt0 = (nt[0] if nt.dim() == 3 else nt).reshape(-1)
    inflow_mask = (nt0 == 1)[:, None].repeat(1, S)              # strategies.jl:177-178
    alg = DeviceAlgebra()
    slots = strategy.solargs.get("stage_workspaces", True)
    if isinstance(strategy, MultipleShooting):
        n_int = len(shooting_ranges(len(time_steps(strategy.tstart, strategy.dt, strategy.tstop)),
                                    strategy.interval_size))
        owned = shard_intervals(n_int, strategy.rank, strategy.world)
        rhs = DeviceRhs(mgn, mgn.ps, fields, target_fields, target_dict, inputs, node_type, edge_feats, senders,
                        receivers, val_mask, inflow_mask, gt, max(len(owned), 1), alg)
        gs, loss, _ = multiple_shooting_step(rhs, alg, mgn.ps, gt, val_mask, strategy.tstart, strategy.dt,
                                             strategy.tstop, strategy.interval_size, strategy.continuity_term,
                                             strategy.solver, strategy.n_sub, owned, slots)
    else:
        rhs = DeviceRhs(mgn, mgn.ps, fields, target_fields, target_dict, inputs, node_type, edge_feats, senders,
                        receivers, val_mask, inflow_mask, gt, 1, alg)
        gs, loss, _ = solver_training_step(rhs, alg, mgn.ps, gt, val_mask, strategy.tstart, strategy.dt,
                                           strategy.tstop, strategy.solver, strategy.n_sub, slots)
    return (gs,), loss


def init_train_step(strategy, t, ta=None):
    """src/strategies.jl:395-415: target = o_norm[f]((data["target|f"][t] - data[f][t]) / dt) for
    each target field (the online normaliser accumulates here), then build_graph.  Solver strategies:
    src/strategies.jl:146-172."""
    if isinstance(strategy, SolverStrategy):
        return _init_solver_step(t)
    mgn, data, meta, fields, target_fields, node_type, edge_feats, senders, receivers, datapoint, mask, _ = t
    cols = []
    for f in target_fields:
        cur, nxt = data[f][datapoint - 1], data["target|" + f][datapoint - 1]
        if isinstance(meta["dt"], (list, tuple)):
            dt = meta["dt"][datapoint] - meta["dt"][datapoint - 1]
        else:
            dt = float(meta["dt"])
        cols.append(mgn.o_norm[f]((nxt - cur) / dt))
    target = torch.cat(cols, dim=1) if len(cols) > 1 else cols[0]
    graph = build_graph(mgn, data, fields, datapoint, node_type, edge_feats, senders, receivers)
    return mgn, graph, target, mask


def train_step(strategy, t):
    """src/strategies.jl:417-422 (derivative) and :174-199 (solver strategies) -> (gs, loss), gs a 1-tuple."""
    if isinstance(strategy, SolverStrategy):
        return _solver_train_step(strategy, t)
    mgn, graph, target, mask = t
    return step_(mgn, graph, target, mask, mse_reduce)
