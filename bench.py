#!/usr/bin/env python
"""bench.py - the reference's headline workload on B200 (BASELINE.json): derivative-training steps
of MeshGraphNets (15 MP steps, latent 128) on a CylinderFlow-shaped synthetic mesh (N=1885,
E=10936), through the product's public API (build_graph -> step! -> Optimisers.update mirrors).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mode bf16|fp32] [--batch B]

One JSON line on rank 0.  metric = MP-step edges/sec inside a full train step
(E x mps x graphs_per_step x n_gpus / step time); `train_steps_per_sec` rides along.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner on fd 1) must not get in
# the way: fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()

MPS, LATENT, HIDDEN = 15, 128, 2
NX, NY = 65, 29            # N = 1885, E = 10936
T_FRAMES = 64              # synthetic trajectory frames resident per rank
METRIC = "mp_step_edges_per_sec_train"
UNIT = "edges/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("MGN_BENCH_MODE", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("MGN_BENCH_BATCH", "64")),
                    help="time windows (graphs) per step per GPU; the reference is batch 1 (reported as `batch1`)")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step as a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--dp", default="fused", choices=["fused", "plain"],
                    help="fused: mgn_backward_dp (bucketed NCCL all-reduce + Adam on a side stream, overlapped with the "
                         "backward pass, library transport); plain: one torch.distributed all-reduce after backward")
    ap.add_argument("--buckets", type=int, default=6)
    ap.add_argument("--no-partition-leg", action="store_true",
                    help="N > 1: skip the graph-partitioned (BASELINE configs[4]) extra measurement")
    ap.add_argument("--partition-edges-per-rank", type=float, default=3.4e6)
    ap.add_argument("--no-shooting-leg", action="store_true",
                    help="skip the interval-sharded MultipleShooting extra measurement (strong scaling over N)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# synthetic workload (shared by both arms)
# ---------------------------------------------------------------------------------------------
def make_workload(batch, seed=1234):
    import mgn_pkg                         # product-side generators: the GPU arm never touches oracle/
    wl = mgn_pkg.pkg
    pos, cells, nt = wl.cylinder_flow_mesh(NX, NY)
    vel = wl.synthetic_velocity(pos, T_FRAMES + 1, seed=seed)
    N = pos.shape[0]
    # `batch` time windows of one trajectory share the topology: block-diagonal graph
    data = {"node_type": np.tile(nt, batch).reshape(1, -1, 1),
            "mesh_pos": np.tile(pos, (batch, 1))[None],
            "cells": np.concatenate([cells + b * N for b in range(batch)], axis=0)[None]}
    return data, vel, nt, N


def sample_frames(vel, step, batch):
    """Window b of step s uses frame (s*batch + b) mod T; returns ([B*N,2] current, [B*N,2] next)."""
    idx = [(step * batch + b) % T_FRAMES for b in range(batch)]
    cur = np.concatenate([vel[i] for i in idx], axis=0)
    nxt = np.concatenate([vel[i + 1] for i in idx], axis=0)
    return cur, nxt


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                mx = max(mx, float(row[1]))
            except Exception:
                continue
            for n, v in zip(names, row[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: torch-CPU restatement of the reference op sequence
# ---------------------------------------------------------------------------------------------
def cpu_reference(batch, seconds, max_steps=None, warmup=1):
    """Times derivative-training steps of oracle/torch_cpu_ref.py (fp32, all host threads) on the
    same CylinderFlow-shaped workload.  Returns (edges/s, steps/s, cores, n_timed)."""
    import mgn_oracle as orc
    import torch_cpu_ref as tref
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would make the baseline
    # single threaded)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    data, vel, nt, N = make_workload(batch)
    cells = data["cells"][0]
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    ef = torch.from_numpy(orc.edge_features(data["mesh_pos"][0], s, r))
    onehot = torch.from_numpy(orc.one_hot(data["node_type"].reshape(-1), 7, 1))
    cfg = orc.ModelConfig(9, 3, 2, LATENT, MPS, HIDDEN)
    p = torch.from_numpy(orc.init_params(cfg))
    opt = {"m": torch.zeros_like(p), "v": torch.zeros_like(p), "t": 0}
    s0 = torch.from_numpy(s.astype(np.int64) - 1)
    r0 = torch.from_numpy(r.astype(np.int64) - 1)
    m0 = torch.from_numpy(orc.node_mask(data["node_type"].reshape(-1), [0, 5]).astype(np.int64) - 1)
    cores = torch.get_num_threads()

    def one(step):
        cur, nxt = sample_frames(vel, step, batch)
        cur, nxt = torch.from_numpy(cur), torch.from_numpy(nxt)
        nf = torch.cat([(cur - cur.mean(0)) / cur.std(0), onehot], dim=1)  # normalise + vcat (graph.jl:75-97)
        tgt = (nxt - cur) / 0.01
        tgt = (tgt - tgt.mean(0)) / tgt.std(0)
        efn = (ef - ef.mean(0)) / ef.std(0).clamp_min(1e-8)
        return tref.train_step(cfg, p, opt, nf, efn, s0, r0, tgt, m0)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    n = 0
    while True:
        one(warmup + n)
        n += 1
        el = time.perf_counter() - t0
        if (max_steps and n >= max_steps) or el >= seconds:
            break
    el = time.perf_counter() - t0
    E = s.shape[0]
    return E * MPS * n / el, batch * n / el, cores, n, el


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (torch-CPU port, all host threads), K
    timed steps of the batch-1 training step after W warm-ups, capped at ~150 s of wall clock.  Rank 0 only."""
    if rank != 0:
        return
    edges_s, steps_s, cores, n, el = cpu_reference(1, 150.0, max_steps=max(1, args.steps),
                                                   warmup=max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": edges_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": n, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": 1000.0 * el / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "train_steps_per_sec": steps_s,
        "config": {"workload": "cylinder_flow_train_step", "nodes": NX * NY, "edges": 10936, "latent": LATENT,
                   "mps": MPS, "graphs_per_step": 1, "note": "reference is batch 1 (batchsize not implemented, "
                   "src/MeshGraphNets.jl:224); Julia/GraphNetCore absent -> torch-CPU port of the same op sequence"},
        "cpu_baseline": {"value": edges_s, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} derivative-training steps (fwd+bwd+Adam), batch 1, {el:.1f} s"},
        "e2e": {"value": edges_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measure(args, rank, world, local_rank, B, steps, warmup, extras):
    """Times `steps` training steps of B time windows per GPU (after `warmup`); returns the JSON line (rank 0)."""
    import mgn_pkg
    pkg = mgn_pkg.pkg

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    mode = pkg.COMPUTE_BF16 if args.mode == "bf16" else pkg.COMPUTE_FP32
    data_h, vel, nt, N = make_workload(B)
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    E = int(senders.shape[0])
    model, ps, st = pkg.build_model(2 + 7, 2, 2, MPS, LATENT, HIDDEN, device=dev, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3, dev),
                           {"velocity": pkg.NormaliserOnline(2, dev), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"velocity": pkg.NormaliserOnline(2, dev)})
    mask = torch.from_numpy(pkg.node_mask(data_h["node_type"].reshape(-1), [0, 5])).to(dev)
    opt = pkg.Adam(1e-4)
    opt_state = opt.setup(mgn.ps)
    strat = pkg.DerivativeTraining()
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}}, "target_features": ["velocity"]}

    # resident synthetic frames: rank-specific windows
    frames = [sample_frames(vel, s + 1000 * rank, B) for s in range(T_FRAMES // max(B, 1) + 2)]
    d_cur = [torch.from_numpy(c).to(dev) for c, _ in frames]
    d_nxt = [torch.from_numpy(n).to(dev) for _, n in frames]
    h_cur = [torch.from_numpy(c).pin_memory() for c, _ in frames]
    h_nxt = [torch.from_numpy(n).pin_memory() for _, n in frames]
    # static device buffers that one step reads (so that the step can be captured in a CUDA graph)
    s_cur = torch.empty_like(d_cur[0])
    s_nxt = torch.empty_like(d_nxt[0])
    data = {"velocity": s_cur[None], "target|velocity": s_nxt[None]}
    loss_buf = torch.zeros(1, device=dev)
    h_loss = torch.zeros(1).pin_memory()
    distributed = world > 1
    comm = None
    if distributed:
        import torch.distributed as dist
        comm = extras if isinstance(extras, pkg.Communicator) else None
    # online-normaliser statistics of all three normalisers in ONE flat buffer: every rank accumulates its own windows,
    # then new = prev + sum_r (state_r - prev) ("online-normaliser stats must be summed across ranks too")
    norms = [mgn.e_norm, mgn.n_norm["velocity"], mgn.o_norm["velocity"]]
    norm_flat = pkg.pack_normaliser_states(norms)
    norm_prev = torch.empty_like(norm_flat)
    fused = args.dp == "fused"

    def step_body(collective=True):
        sync = distributed and collective
        if sync:
            norm_prev.copy_(norm_flat)                  # the common statistics before this step's accumulation
        t = pkg.init_train_step(strat, (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders,
                                        receivers, 1, mask, None))
        if fused:
            # forward + loss + backward whose finished gradient buckets are all-reduced (mean) and fed to Adam on a
            # side stream while the rest of the backward pass runs (mgn_backward_dp)
            _, graph, target, _ = t
            gs, loss = pkg.step_dp_(mgn, graph, target, mask, opt=opt, opt_state=opt_state,
                                    comm=comm if sync else None, n_buckets=args.buckets)
        else:
            gs, loss = pkg.train_step(strat, t)
            for g in gs:
                if sync:
                    pkg.allreduce_mean_(g, world)      # one NCCL all-reduce of the flat fp32 gradient
                opt.update(opt_state, mgn.ps, g)
        if sync:                                         # new = prev + sum over ranks of (state_r - prev): one collective
            if comm is not None:
                comm.allreduce_normaliser_(norm_flat, norm_prev)
            else:
                pkg.allreduce_normaliser_(norm_flat, norm_prev)
        loss_buf.copy_(loss)

    # ---- optional CUDA graph of the step (all library calls only enqueue on the current stream)
    graph = None
    use_graph = not args.no_graph          # NCCL collectives are capturable: the DP step is one graph too
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(2):  # warm the workspace / caches outside capture
            s_cur.copy_(d_cur[i % len(d_cur)]); s_nxt.copy_(d_nxt[i % len(d_nxt)])
            step_body()
    torch.cuda.synchronize()
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            t_state = opt_state["t"]
            with torch.cuda.graph(graph):
                step_body()
            opt_state["t"] = t_state  # NOTE: Adam bias correction is frozen inside the graph replay
        except Exception as e:  # capture unsupported -> plain launches
            print(f"[bench] CUDA graph capture failed ({e}); using plain launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 256 MB > 126 MB L2

    def run_step(i, host_inputs):
        k = i % len(d_cur)
        if host_inputs:
            s_cur.copy_(h_cur[k], non_blocking=True)
            s_nxt.copy_(h_nxt[k], non_blocking=True)
        else:
            s_cur.copy_(d_cur[k]); s_nxt.copy_(d_nxt[k])
        if graph is not None:
            graph.replay()
        else:
            step_body()
        if host_inputs:
            h_loss.copy_(loss_buf, non_blocking=True)

    def timed(K, host_inputs):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        for i in range(K):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            evs[i][0].record()
            run_step(i, host_inputs)
            evs[i][1].record()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        return sum(ms) / K

    for i in range(warmup):
        run_step(i, False)
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        ms_dev = timed(steps, False)
        ms_e2e = timed(steps, True)
    clocks = clk.summary()
    if distributed:
        tt = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(tt[0]), float(tt[1])
    final_loss = float(loss_buf.cpu())

    if rank != 0:
        return None
    edges_per_step = E * MPS * world            # E already includes the batch (block-diagonal graph)
    value = edges_per_step / (ms_dev * 1e-3)
    e2e = edges_per_step / (ms_e2e * 1e-3)
    h2d = int(s_cur.numel() * 4 + s_nxt.numel() * 4)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if mode == pkg.COMPUTE_BF16 else "f32", "data": "synthetic",
        "train_steps_per_sec": B * world / (ms_dev * 1e-3),
        "config": {"workload": "cylinder_flow_train_step", "nodes": NX * NY, "edges": E // B, "latent": LATENT,
                   "mps": MPS, "hidden_layers": HIDDEN, "graphs_per_step_per_gpu": B, "parallelism": f"dp{world}",
                   "compute_mode": args.mode, "cuda_graph": graph is not None,
                   "dp": (f"mgn_backward_dp: {args.buckets} gradient buckets, NCCL avg + Adam per bucket on a side stream, "
                          "overlapped with the backward pass; normaliser statistics merged across ranks every step"
                          if fused else "one all-reduce after backward, then Adam") if world > 1 else
                         ("mgn_backward_dp: Adam per gradient bucket on a side stream" if fused else "step! then Adam"),
                   "l2": "flushed between timed steps (256 MB write); per-step CUDA events, max over ranks"},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "clocks": clocks, "final_loss": final_loss,
    }
    def plain_step():  # rank 0 only (kernel-family timing): no collective, the other ranks have left
        s_cur.copy_(d_cur[0]); s_nxt.copy_(d_nxt[0])
        step_body(collective=False)
    if extras:
        line.update(extra_measurements(args, pkg, model, mgn, E, B, dev, plain_step, N * B))
        line["gpu_launches"] = line.get("gpu_launches_per_step", 0) * steps
    return line


def run_ours(args, rank, world, local_rank):
    import mgn_pkg
    pkg = mgn_pkg.pkg
    extras = True
    if world > 1 and args.dp == "fused":
        torch.cuda.set_device(local_rank)
        extras = pkg.Communicator.from_torch_distributed(torch.device("cuda", local_rank))   # the library's own NCCL comm
    line = measure(args, rank, world, local_rank, args.batch, args.steps, args.warmup, extras)
    if world > 1 and not args.no_partition_leg:
        try:
            leg = partition_leg(args, pkg, rank, world, local_rank, extras if isinstance(extras, pkg.Communicator) else None)
        except Exception as e:       # never lose the main line
            leg = {"error": repr(e)}
        if line is not None:
            line["partitioned_mesh"] = leg
    if not args.no_shooting_leg:
        try:
            leg = shooting_leg(args, pkg, rank, world, local_rank, extras if isinstance(extras, pkg.Communicator) else None)
        except Exception as e:
            leg = {"error": repr(e)}
        if line is not None:
            line["multiple_shooting_sharded"] = leg
    if world == 1 and args.batch != 1:
        # the reference's own granularity: ONE window per step (batchsize is "not implemented yet",
        # src/MeshGraphNets.jl:224) - a latency number, reported beside the throughput headline
        b1 = measure(args, rank, world, local_rank, 1, min(args.steps, 20), max(3, min(args.warmup, 5)), False)
        if line is not None and b1 is not None:
            from bench_roofline import peaks, survey_flops
            pk = peaks()
            tf = survey_flops(10936, NX * NY, LATENT, HIDDEN + 2, MPS, 9, 3, 2) * 3 / (b1["ms_per_step"] * 1e-3) / 1e12
            line["batch1"] = {"ms_per_step": b1["ms_per_step"], "value": b1["value"], "unit": UNIT,
                              "train_steps_per_sec": b1["train_steps_per_sec"], "e2e_ms_per_step": b1["e2e"]["ms_per_step"],
                              "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops_sustained"],
                                           "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops_sustained"],
                                           "note": "one window = 86 edge tiles < 148 SMs and a 7 MB working set in L2: "
                                                   "launch / latency bound; algorithmic FLOPs of SURVEY 8d (3 x forward) "
                                                   "against the sustained bf16 peak (84 us per step at peak)"}}
    if line is not None:
        emit(line)


def partition_leg(args, pkg, rank, world, local_rank, comm):
    """BASELINE configs[4] on the driver's record: one training step of a Kuhn tetrahedral grid partitioned over the
    ranks (about --partition-edges-per-rank edges each; 126^3 nodes / 27.6 M edges at N = 8), halo rows exchanged
    every message-passing step through mgn_halo_exchange (grouped NCCL send/recv on the compute stream), the parameter
    gradient summed with mgn_dp_allreduce.  Device-timed, max over ranks; exchange time = sum of CUDA-event pairs
    around the exchanges of rank 0's stream inside the same steps."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    n = max(8, int(round((args.partition_edges_per_rank * world / 13.8) ** (1.0 / 3.0))))
    N = n ** 3
    s, r = pkg.parse_edges(pkg.tet_grid_edges(n))
    E = int(s.shape[0])
    part = pkg.build_partition_rank(N, s, r, world, rank)
    del s, r
    rng = np.random.default_rng(1234 + rank)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nf = to(rng.normal(size=(part.n_local, 4)).astype(np.float32))
    ef = to(rng.normal(size=(len(part.edge_ids), 4)).astype(np.float32))
    tgt = to(rng.normal(size=(part.n_local, 3)).astype(np.float32))
    mask = to(np.arange(1, part.n_own + 1, dtype=np.int32))
    mode = pkg.COMPUTE_BF16 if args.mode == "bf16" else pkg.COMPUTE_FP32
    model, ps, _ = pkg.build_model(4, 3, 3, MPS, LATENT, HIDDEN, device=dev, compute_mode=mode)
    pm = pkg.PartitionedModel(model, part, nf, ef, device=dev)
    base = pkg.AbiExchange(part, world, comm, model) if comm is not None else pkg.DistExchange(part, world, dev, model)
    ev = []

    def timed_exchange(sends, direction):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = base(sends, direction)
        b.record()
        ev.append((a, b))
        return out

    def step(ex):
        grads, losses, _ = pkg.run_partitioned_step([pm], ps, [tgt], [mask], N, ex, pkg.masked_mse_partial)
        if comm is not None:
            comm.allreduce_sum_(grads[0]); comm.allreduce_sum_(losses[0])
        else:
            dist.all_reduce(grads[0]); dist.all_reduce(losses[0])
        return losses[0]

    step(base)
    torch.cuda.synchronize()
    dist.barrier()
    K = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(timed_exchange)
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / K
    ex_ms = sum(a.elapsed_time(b) for a, b in ev) / K
    # the same event pairs with the ranks aligned on the device right before every exchange (a one-element all-reduce
    # absorbs the skew): what is left is the transfer itself - concatenation, grouped send / recv - without the wait for
    # the slowest peer that the figure above includes
    ev_sync, skew = [], torch.zeros(1, device=dev)

    def aligned_exchange(sends, direction):
        if comm is not None:
            comm.allreduce_sum_(skew)
        else:
            dist.all_reduce(skew)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = base(sends, direction)
        b.record()
        ev_sync.append((a, b))
        return out

    step(aligned_exchange)
    torch.cuda.synchronize()
    ex_sync_ms = sum(a.elapsed_time(b) for a, b in ev_sync)
    # the same step - 33 stages, 29 halo exchanges, 2 all-reduces - captured once and replayed as ONE CUDA graph
    ms, captured = ms_eager, False
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(base)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss_g = step(base)
        graph.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(K):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms, captured, loss = e0.elapsed_time(e1) / K, True, loss_g
    except Exception as e:       # capture not possible on this stack: the eager number stands
        print(f"[bench] partitioned step: CUDA graph capture failed ({e!r}); eager launches", file=sys.stderr)
        torch.cuda.synchronize()
    halo = sum(len(v) for v in part.recv_rows.values())
    tt = torch.tensor([ms, ex_ms, float(halo), float(len(part.edge_ids)), ms_eager, ex_sync_ms], device=dev,
                      dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ws_gb = model.workspace_bytes(pm.graph.index, True) / 1e9
    lat_b, grad_b = model.halo_row_bytes(pkg.HALO_LATENT), model.halo_row_bytes(pkg.HALO_GRAD)
    return {"workload": f"kuhn_tet_grid_{n}^3_partitioned_train_step", "nodes": N, "edges": E, "mps": MPS,
            "ms_per_step": float(tt[0]), "mp_step_edges_per_sec": E * MPS / (float(tt[0]) * 1e-3),
            "cuda_graph": captured, "ms_per_step_eager": float(tt[4]),
            "exchange_ms_per_step_max_rank": float(tt[1]), "exchange_ms_per_step_ranks_aligned": float(tt[5]),
            "exchanges_per_step": 2 * MPS - 1,
            "halo_rows_max_rank": int(tt[2]), "edges_max_rank": int(tt[3]),
            "halo_bytes_per_step_max_rank": int(tt[2]) * ((MPS - 1) * lat_b + MPS * grad_b),
            "workspace_gb_rank0": ws_gb, "steps": K,
            "transport": "mgn_halo_exchange (library NCCL, grouped send/recv on the compute stream)" if comm is not None
                         else "torch.distributed all_to_all_single",
            "final_loss": float(loss.cpu())}


def shooting_leg(args, pkg, rank, world, local_rank, comm):
    """SURVEY 8e row 4 on the driver's record: one MultipleShooting training step (src/strategies.jl:310-386) over a
    200-observation CylinderFlow trajectory = 40 shooting intervals (interval_size 6, Euler, 15 MP steps), the intervals
    sharded over the ranks (rank r integrates intervals r, r + world, ... in lock-step as one block-diagonal graph),
    loss and gradient SUM-all-reduced.  Strong scaling: the work is fixed, so `ms_per_step` falls with N."""
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    T = 200
    pos, cells, nt = pkg.cylinder_flow_mesh(NX, NY)
    vel = pkg.synthetic_velocity(pos, T + 1, seed=1)
    data_h = {"node_type": nt.reshape(1, -1, 1), "mesh_pos": pos[None], "cells": cells[None]}
    node_type, senders, receivers, ef = pkg.create_base_graph(data_h, 6, 0, device=dev)
    mode = pkg.COMPUTE_BF16 if args.mode == "bf16" else pkg.COMPUTE_FP32
    model, ps, st = pkg.build_model(9, 2, 2, MPS, LATENT, HIDDEN, device=dev, compute_mode=mode)
    mgn = pkg.GraphNetwork(model, ps, st, pkg.NormaliserOnline(3, dev),
                           {"velocity": pkg.NormaliserOnline(2, dev), "node_type": pkg.NormaliserOfflineMinMax(0.0, 1.0)},
                           {"velocity": pkg.NormaliserOnline(2, dev)})
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    data = {"velocity": to(vel[:T]), "target|velocity": to(vel[1:T + 1]),
            "node_type": to(nt.reshape(1, -1, 1).astype(np.int32))}
    for f in range(3):
        mgn.n_norm["velocity"](data["velocity"][f])
        mgn.o_norm["velocity"]((data["velocity"][f + 1] - data["velocity"][f]) / 0.01)
    mgn.e_norm(ef)
    meta = {"dt": 0.01, "features": {"velocity": {"dim": 2}}, "target_features": ["velocity"]}
    vm = to(pkg.val_mask(nt, [0, 5], 2))
    t = (mgn, data, meta, ["velocity"], ["velocity"], node_type, ef, senders, receivers, 1, None, vm)
    n_int = len(pkg.shooting_ranges(T, 6))
    strat = pkg.MultipleShooting(0.0, 0.01, 0.01 * (T - 1), "euler", interval_size=6, continuity_term=100, rank=rank,
                                 world=world)
    tt = pkg.init_train_step(strat, t, None)

    def step():
        (gs,), loss = pkg.train_step(strat, tt)
        if world > 1:
            if comm is not None:
                comm.allreduce_sum_(gs); comm.allreduce_sum_(loss)
            else:
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
    ms = float(np.median(times))
    E = int(senders.shape[0])
    return {"workload": "cylinder_flow_multiple_shooting_step_intervals_sharded", "observations": T, "intervals": n_int,
            "intervals_per_rank": len(pkg.shard_intervals(n_int, 0, world)), "solver": "euler", "scaling": "strong",
            "ms_per_step": ms, "mp_step_edges_per_sec_train_equiv": (5 + 5 / 3) * n_int * E * MPS / (ms * 1e-3),
            "loss": float(loss.cpu()), "grad_norm": float(gs.norm().cpu())}


def extra_measurements(args, pkg, model, mgn, E, B, dev, step_fn, n_nodes):
    """gpu_launches, roofline of the dominant kernel, and the CPU baseline (rank 0, N=1 only)."""
    out = {}
    try:
        from bench_roofline import roofline_and_launches
        out.update(roofline_and_launches(args, pkg, model, mgn, E, B, dev, step_fn, n_nodes))
    except Exception as e:  # never lose the main line
        out["roofline_error"] = repr(e)
    if args.gpus == 1:
        try:
            edges_s, steps_s, cores, n, el = cpu_reference(1, args.cpu_seconds)
            out["cpu_baseline"] = {"value": edges_s, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{n} batch-1 derivative-training steps (fwd+bwd+Adam) of the torch-CPU "
                                             f"port, {el:.1f} s", "train_steps_per_sec": steps_s}
        except Exception as e:
            out["cpu_baseline_error"] = repr(e)
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
