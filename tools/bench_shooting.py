#!/usr/bin/env python
"""Solver-strategy training step on one B200 (src/strategies.jl MultipleShooting / SolverTraining through the host
mirror): CylinderFlow-shaped mesh, 15 MP steps, tsteps = 0:0.01:0.49 (50 observations), interval_size 6 -> 10 shooting
intervals.  Times one train_step (lock-step solve + reverse sweep) with CUDA events and, for comparison, the same step
with the intervals solved one after the other (the reference's order, strategies.jl:349-362) through the same
kernels.  Prints one JSON line per configuration.

Under torchrun (WORLD_SIZE > 1, one process per GPU) the shooting intervals of a longer trajectory (`--obs`, default 200
observations -> 40 intervals) are sharded over the ranks (shooting.shard_intervals), loss and gradient are SUM-all-reduced
over NCCL, and the time is the max over ranks:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 \
        tools/bench_shooting.py bf16 --sharded"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import mgn_pkg  # noqa: E402

pkg = mgn_pkg.pkg


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def sharded(mode):
    """Interval-sharded MultipleShooting step over the ranks of a torchrun job."""
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = int(sys.argv[sys.argv.index("--obs") + 1]) if "--obs" in sys.argv else 200
    pos, cells, nt = pkg.cylinder_flow_mesh(65, 29)
    vel = pkg.synthetic_velocity(pos, T + 1, seed=1)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    model, ps, st = pkg.build_model(9, 2, 2, 15, 128, 2, device=dev, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3, dev),
                           {"velocity": pkg.NormaliserOnline(2, dev), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"velocity": pkg.NormaliserOnline(2, dev)})
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    data = {"velocity": to(vel[:T]), "target|velocity": to(vel[1:T + 1]), "node_type": to(nt.reshape(1, -1, 1).astype(np.int32))}
    for f in range(3):
        mgn.n_norm["velocity"](data["velocity"][f])
        mgn.o_norm["velocity"]((data["velocity"][f + 1] - data["velocity"][f]) / 0.01)
    mgn.e_norm(ef)
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}}, "target_features": ["velocity"]}
    vm = to(pkg.val_mask(nt, [0, 5], 2))
    t = (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders, receivers, 1, None, vm)
    tstop = 0.01 * (T - 1)
    n_int = len(pkg.shooting_ranges(T, 6))
    strat = pkg.MultipleShooting(0.0, 0.01, tstop, "euler", interval_size=6, continuity_term=100, rank=rank, world=world)
    tt = pkg.init_train_step(strat, t, None)

    def step():
        (gs,), loss = pkg.train_step(strat, tt)
        pkg.allreduce_sum_(gs, loss)
        return gs, loss
    step()
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gs, loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms.cpu()))
    if rank == 0:
        ms = float(np.median(times))
        E = int(senders.shape[0])
        print(json.dumps({"workload": "cylinder_flow_multiple_shooting_step_sharded", "n_gpus": world, "observations": T,
                          "intervals": n_int, "intervals_per_rank": len(pkg.shard_intervals(n_int, 0, world)),
                          "solver": "euler", "mode": "bf16" if mode == pkg.COMPUTE_BF16 else "fp32",
                          "ms_per_train_step": ms, "scaling": "strong",
                          "mp_step_edges_per_sec_train_equiv": (5 + 5 / 3) * n_int * E * 15 / (ms * 1e-3),
                          "loss": float(loss.cpu()), "grad_norm": float(gs.norm().cpu())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def chain(mode):
    """BASELINE configs[3]: 1-D chain of 100 000 nodes (src/dataset.jl:379-382 through parse_edges, E = 199 998,
    F_e = 2), one target field `u` and one non-target input field, 15 MP steps, MultipleShooting with 26 observations
    and interval_size 6 -> 5 intervals in lock-step = one 500 000-node / 999 990-edge block-diagonal graph, Euler."""
    dev = torch.device("cuda", 0)
    n, T = 100_000, 26
    rng = np.random.default_rng(11)
    pos = np.linspace(0.0, 1.0, n, dtype=np.float32)[:, None]
    nt = np.zeros(n, dtype=np.int32)
    nt[0], nt[-1] = 4, 5
    xs = pos[:, 0]
    u = np.stack([np.sin(2 * np.pi * (xs - 0.02 * k)) + 0.05 * rng.normal(size=n) for k in range(T)])[:, :, None]
    load = np.cos(4 * np.pi * xs)[None, :, None].repeat(T, 0)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "edges": pkg.chain_edges(n)}
    meta = {"dt": 0.01, "features": {"u": {"dim": 1}, "load": {"dim": 1}}, "target_features": ["u"]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    data = {"u": to(u.astype(np.float32)), "load": to(load.astype(np.float32)),
            "node_type": to(nt.reshape(1, -1, 1).astype(np.int32))}
    vm = to(pkg.val_mask(nt, [0], 1))
    model, ps, st = pkg.build_model(1 + 1 + 7, 1, 1, 15, 128, 2, device=dev, compute_mode=mode, seed=5)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOfflineMeanStd(0.0, 1e-5),
                           {"u": pkg.NormaliserOfflineMeanStd(0.0, 0.7), "load": pkg.NormaliserOfflineMinMax(-1.0, 1.0),
                            "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"u": pkg.NormaliserOfflineMeanStd(0.0, 3.0)})
    strat = pkg.MultipleShooting(0.0, 0.01, 0.25, "euler", interval_size=6, continuity_term=100)
    t = (mgn, data, meta, ["u", "load"], ["u"], node_type, ef, senders, receivers, 1, None, vm)
    tt = pkg.init_train_step(strat, t, None)
    n_int = len(pkg.shooting_ranges(T, 6))
    ms, ((gs,), loss) = timed(lambda: pkg.train_step(strat, tt), 2)
    E = int(senders.shape[0])
    print(json.dumps({"workload": "chain_100k_multiple_shooting_step", "nodes": n, "edges": E, "mps": 15,
                      "mode": "bf16" if mode == pkg.COMPUTE_BF16 else "fp32", "solver": "euler", "observations": T,
                      "intervals": n_int, "lockstep_graph_nodes": n * n_int, "lockstep_graph_edges": E * n_int,
                      "ms_per_train_step": ms,
                      "mp_step_edges_per_sec_train_equiv": (5 + 5 / 3) * n_int * E * 15 / (ms * 1e-3),
                      "loss": float(loss.cpu()), "grad_finite": bool(torch.isfinite(gs).all()),
                      "hbm_gb_allocated": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def main():
    mode = pkg.COMPUTE_BF16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else pkg.COMPUTE_FP32
    if "--sharded" in sys.argv:
        return sharded(mode)
    if "--chain" in sys.argv:
        return chain(mode)
    dev = torch.device("cuda", 0)
    pos, cells, nt = pkg.cylinder_flow_mesh(65, 29)
    T = 50
    vel = pkg.synthetic_velocity(pos, T + 1, seed=1)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    N, E = pos.shape[0], int(senders.shape[0])
    model, ps, st = pkg.build_model(9, 2, 2, 15, 128, 2, device=dev, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3, dev),
                           {"velocity": pkg.NormaliserOnline(2, dev), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"velocity": pkg.NormaliserOnline(2, dev)})
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    data = {"velocity": to(vel[:T]), "target|velocity": to(vel[1:T + 1]), "node_type": to(nt.reshape(1, -1, 1).astype(np.int32))}
    for f in range(3):
        mgn.n_norm["velocity"](data["velocity"][f])
        mgn.o_norm["velocity"]((data["velocity"][f + 1] - data["velocity"][f]) / 0.01)
    mgn.e_norm(ef)
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}}, "target_features": ["velocity"]}
    vm = to(pkg.val_mask(nt, [0, 5], 2))
    t = (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders, receivers, 1, None, vm)
    n_int = len(pkg.shooting_ranges(50, 6))
    for solver in ("euler", "tsit5"):
        stages = len(pkg.RK_TABLEAUS[solver][2])
        strat = pkg.MultipleShooting(0.0, 0.01, 0.49, solver, interval_size=6, continuity_term=100)
        tt = pkg.init_train_step(strat, t, None)
        ms, ((gs,), loss) = timed(lambda: pkg.train_step(strat, tt), 3)

        def sequential():
            g, l = None, 0.0
            for i in range(n_int):
                s = pkg.MultipleShooting(0.0, 0.01, 0.49, solver, interval_size=6, continuity_term=100, rank=i, world=n_int)
                (gi,), li = pkg.train_step(s, tt)
                g = gi if g is None else g + gi
                l = l + li
            return (g,), l
        ms_seq, ((gs_seq,), loss_seq) = timed(sequential, 1)
        # RHS evaluations of one step: forward sweep (5 steps x stages) + reverse sweep (5 x stages fwd+bwd), K intervals each
        rhs_fwd, rhs_bwd = 2 * 5 * stages, 5 * stages
        print(json.dumps({
            "workload": "cylinder_flow_multiple_shooting_step", "nodes": N, "edges": E, "mps": 15,
            "mode": "bf16" if mode == pkg.COMPUTE_BF16 else "fp32", "solver": solver, "intervals": n_int,
            "interval_size": 6, "ms_per_train_step_lockstep": ms, "ms_per_train_step_sequential": ms_seq,
            "speedup_lockstep": ms_seq / ms, "rhs_forward_evals": rhs_fwd, "rhs_backward_evals": rhs_bwd,
            # in units of one forward+backward pass (a forward-only evaluation counts 1/3, the FLOP convention of bench.py)
            "mp_step_edges_per_sec_train_equiv": (rhs_bwd + (rhs_fwd - rhs_bwd) / 3) * n_int * E * 15 / (ms * 1e-3),
            "loss": float(loss.cpu()), "loss_sequential": float(loss_seq.cpu()),
            "grad_rel_diff_vs_sequential": float(((gs - gs_seq).norm() / gs_seq.norm()).cpu())}), flush=True)
    strat = pkg.SolverTraining(0.0, 0.01, 0.49, "euler")
    ms, ((gs,), loss) = timed(lambda: pkg.train_step(strat, pkg.init_train_step(strat, t, None)), 2)
    print(json.dumps({"workload": "cylinder_flow_solver_training_step", "solver": "euler", "saves": 50,
                      "mode": "bf16" if mode == pkg.COMPUTE_BF16 else "fp32", "ms_per_train_step": ms,
                      "loss": float(loss.cpu()), "grad_finite": bool(torch.isfinite(gs).all())}), flush=True)


if __name__ == "__main__":
    main()
