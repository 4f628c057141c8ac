#!/usr/bin/env python
"""BASELINE config 3: NeuralODE rollout of a CylinderFlow-shaped model through the src/solve.jl RHS mirror
(ode_func_eval -> ode_step: inflow overwrite, build_graph normalisers, model forward, inverse_data, val_mask),
50 saved steps with fixed-step Euler (examples/cylinder_flow/cylinder_flow.jl:79-84) and with the 6-stage
Tsit5 step (src/solve.jl:58).  Prints one JSON line: ms per RHS evaluation and per rollout, plain launches and
with the RHS replayed as a CUDA graph."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import mgn_pkg  # noqa: E402

pkg = mgn_pkg.pkg


def main():
    mode = pkg.COMPUTE_BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else pkg.COMPUTE_FP32
    dev = torch.device("cuda", 0)
    pos, cells, nt = pkg.cylinder_flow_mesh(65, 29)
    T = 51
    vel = pkg.synthetic_velocity(pos, T, seed=1)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    N, E = pos.shape[0], int(senders.shape[0])
    model, ps, st = pkg.build_model(9, 2, 2, 15, 128, 2, device=dev, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3, dev),
                           {"velocity": pkg.NormaliserOnline(2, dev), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"velocity": pkg.NormaliserOnline(2, dev)})
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    data = {"velocity": to(vel)}
    for f in range(3):                                   # accumulate some normaliser statistics first
        mgn.n_norm["velocity"](data["velocity"][f])
        mgn.o_norm["velocity"]((data["velocity"][f + 1] - data["velocity"][f]) / 0.01)
    mgn.e_norm(ef)
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}}, "target_features": ["velocity"]}
    val_mask = to(pkg.val_mask(nt, [0, 5], 2))
    inflow = to(np.repeat((nt == 1)[:, None], 2, axis=1))
    saves = np.arange(0, 51, dtype=np.float32) * np.float32(0.01)
    init = {"velocity": data["velocity"][0]}

    def run(solver):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol, ts = pkg.rollout(mgn, init, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders,
                              receivers, val_mask, inflow, data, 0.0, 0.5, 0.01, saves, solver=solver)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, sol

    run("euler")
    ms_euler, sol = run("euler")
    ms_tsit, _ = run("tsit5")
    # one RHS evaluation replayed as a CUDA graph (what a solver loop on the device would cost)
    x = init["velocity"].clone()
    p = (mgn, mgn.ps, {}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders, receivers, val_mask)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            y = pkg.ode_step(x, p, 0.0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = pkg.ode_step(x, p, 0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    # the whole 50-step rollout as ONE CUDA graph (CapturedRollout)
    args = (mgn, init, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders, receivers, val_mask,
            inflow, data, 0.0, 0.5, 0.01, saves)
    captured = {}
    for solver in ("euler", "tsit5"):
        cap = pkg.CapturedRollout(*args, solver=solver)
        cap.replay()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            cap.replay()
        c1.record()
        torch.cuda.synchronize()
        captured[solver] = c0.elapsed_time(c1) / 3
    print(json.dumps({"workload": "cylinder_flow_rollout_50_steps", "euler_50_steps_one_graph_ms": captured["euler"],
                      "tsit5_50_steps_one_graph_ms": captured["tsit5"], "nodes": N, "edges": E, "mps": 15,
                      "mode": "bf16" if mode == pkg.COMPUTE_BF16 else "fp32",
                      "euler_50_steps_ms": ms_euler, "tsit5_50_steps_ms": ms_tsit,
                      "rhs_ms_plain": ms_euler / 50, "rhs_ms_cuda_graph": e0.elapsed_time(e1) / 50,
                      "mp_step_edges_per_sec_inference": E * 15 / (e0.elapsed_time(e1) / 50 * 1e-3),
                      "final_state_finite": bool(torch.isfinite(sol[-1]).all())}), flush=True)


if __name__ == "__main__":
    main()
