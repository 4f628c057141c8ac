"""Mirror of src/solve.jl: ode_step (:188-219), ode_func_eval (:147-158) and rollout (:42-68) with the
fixed-step Euler configuration of examples/cylinder_flow/cylinder_flow.jl:79-84.  The adaptive
Tsit5 driver itself is OrdinaryDiffEq's (out of scope); tsit5_step below restates one explicit
Tsit5 step so that the 6-RHS-evaluations-per-step workload of config 3 can be timed."""
from __future__ import annotations

import numpy as np
import torch

from .graph import build_graph


def ode_step(x, p, t):
    """src/solve.jl:188-219.  x is [N, sum(target dims)]; p = (mgn, ps, inputs, fields, meta,
    target_fields, target_dict, node_type, edge_features, senders, receivers, val_mask)."""
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
    return buf * val_mask


def ode_func_eval(x, p, t):
    """src/solve.jl:147-158: overwrite the inflow nodes from data, then ode_step."""
    (mgn, ps, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
     receivers, val_mask, inflow_mask, saves_dt) = p
    if inflow_mask is not None:
        # floor(Int, t / saves_dt) + 1 with Float32 t and saves_dt (0-based here)
        idx = int(np.floor(np.float32(t) / np.float32(saves_dt)))
        cur = torch.cat([data[f][idx] for f in target_fields], dim=1)
        x.copy_(torch.where(inflow_mask, cur, x))  # in place: the reference mutates the solver state
    return ode_step(x, (mgn, ps, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats,
                        senders, receivers, val_mask), t)


def rollout(mgn, initial_state, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
            receivers, val_mask, inflow_mask, data, start, stop, dt, saves, solver="euler"):
    """src/solve.jl:42-68 with ``solve(prob, Euler(); adaptive=false, dt=dt, saveat=saves)``.
    Returns (list of saved states, times)."""
    x = torch.cat([initial_state[f] for f in target_fields], dim=1).clone()
    inputs = {k: v for k, v in initial_state.items() if k not in target_dict}
    p = (mgn, mgn.ps, data, inputs, fields, meta, target_fields, target_dict, node_type, edge_feats, senders,
         receivers, val_mask, inflow_mask, saves[1] - saves[0])
    sol, ts = [x.clone()], [float(saves[0])]
    step = euler_step if solver == "euler" else tsit5_step
    n_steps = len(saves) - 1
    for i in range(n_steps):
        t = np.float32(saves[i])  # tstops = saves: the integrator lands on the save times
        x = step(lambda xx, tt: ode_func_eval(xx, p, tt), x, t, dt)
        sol.append(x.clone())
        ts.append(float(saves[i + 1]))
    return sol, ts


def euler_step(f, x, t, dt):
    return x + dt * f(x, t)


_TSIT5_C = (0.161, 0.327, 0.9, 0.9800255409045097, 1.0)
_TSIT5_A = (
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
)
_TSIT5_B = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
            2.324710524099774)


def tsit5_step(f, x, t, dt):
    """One explicit Tsitouras 5(4) step (6 RHS evaluations; the 7th FSAL stage is the next step's
    first) - the per-step RHS workload of OrdinaryDiffEq.Tsit5 at src/solve.jl:58."""
    k = [f(x, t)]
    for c, a in zip(_TSIT5_C, _TSIT5_A):
        xi = x
        for aj, kj in zip(a, k):
            xi = xi + (dt * aj) * kj
        k.append(f(xi, t + c * dt))
    out = x
    for bj, kj in zip(_TSIT5_B, k):
        out = out + (dt * bj) * kj
    return out
