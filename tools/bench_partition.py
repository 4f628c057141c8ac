#!/usr/bin/env python
"""BASELINE config 5: a synthetic 3-D tetrahedral (Kuhn-subdivided) grid mesh, graph-partitioned into slabs
over the GPUs of one box with a halo exchange of boundary-node latents at every message-passing step
(NCCL all-to-all over NVLink).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29533 \
        tools/bench_partition.py --grid 126 --steps 3 [--mode bf16]

Prints one JSON line on rank 0: train-step time (max over ranks, CUDA events), MP-step edges/s of the whole
mesh, halo bytes per exchange.  With --check the partitioned loss is compared with the unpartitioned one
(single-process sizes only)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import mgn_pkg  # noqa: E402

pkg = mgn_pkg.pkg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=48)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--mps", type=int, default=15)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.grid
    N = n ** 3
    s, r = pkg.parse_edges(pkg.tet_grid_edges(n))
    E = int(s.shape[0])
    part = pkg.build_partition_rank(N, s, r, world, rank)
    del s, r
    rng = np.random.default_rng(1234 + rank)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nf = to(rng.normal(size=(part.n_local, 4)).astype(np.float32))       # node features + one-hot stand-in
    ef = to(rng.normal(size=(len(part.edge_ids), 4)).astype(np.float32)) # [rel ; |rel|] of a 3-D mesh
    tgt = to(rng.normal(size=(part.n_local, 3)).astype(np.float32))
    mask = to(np.arange(1, part.n_own + 1, dtype=np.int32))
    mode = pkg.COMPUTE_BF16 if args.mode == "bf16" else pkg.COMPUTE_FP32
    model, ps, _ = pkg.build_model(4, 3, 3, args.mps, 128, 2, device=dev, compute_mode=mode)
    pm = pkg.PartitionedModel(model, part, nf, ef, device=dev)
    ex = pkg.DistExchange(part, world, dev, model) if world > 1 else pkg.LocalExchange()
    halo_rows = sum(len(v) for v in part.recv_rows.values())

    def step():
        grads, losses, _ = pkg.run_partitioned_step([pm], ps, [tgt], [mask], N, ex, pkg.masked_mse_partial)
        if world > 1:
            dist.all_reduce(grads[0])
            dist.all_reduce(losses[0])
        return grads[0], losses[0]

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g, loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"tet_grid_{n}^3_partitioned", "nodes": N, "edges": E, "n_gpus": world,
                          "mps": args.mps, "mode": args.mode, "ms_per_train_step": float(ms),
                          "mp_step_edges_per_sec_train": E * args.mps / (float(ms) * 1e-3),
                          "halo_rows_rank0": halo_rows, "halo_bytes_per_exchange_rank0": halo_rows * 256,
                          "exchanges_per_step": 2 * args.mps - 1, "loss": float(loss.cpu()),
                          "mem_gb_rank0": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
