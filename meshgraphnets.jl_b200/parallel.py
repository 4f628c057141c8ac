"""Data-parallel plumbing of the training step: one process per GPU, every rank
owns a strided subset of the (shuffled) time windows of a trajectory (src/MeshGraphNets.jl:364-370
iterates them sequentially; P-way DP is the `batchsize` the reference leaves unimplemented, :224), the
flat Float32 gradient is all-reduced and averaged, and the online-normaliser statistics are summed.
torch.distributed is only the transport (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_windows(n_windows: int, rank: int, world: int):
    """Windows rank, rank + world, ... of the permutation: disjoint, covering, balanced to within one."""
    return list(range(rank, n_windows, world))


def allreduce_mean_(flat_grads: torch.Tensor, world: int | None = None):
    """In-place mean of the flat gradient over all ranks (a single collective: 11.5 MB at CylinderFlow size)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grads
    world = world or dist.get_world_size()
    if world == 1:
        return flat_grads
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
    flat_grads.mul_(1.0 / world)
    return flat_grads


def allreduce_sum_(*tensors):
    """In-place SUM over ranks: loss and gradient of a MultipleShooting step whose intervals are sharded over
    processes (shooting.shard_intervals) - every interval contributes exactly once."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def allreduce_normaliser_(state: torch.Tensor, prev: torch.Tensor):
    """Online-normaliser state [sum | sum_sq | count | num_acc] after a step in which every rank
    accumulated its own window on top of the common `prev`: new = prev + sum_r (state_r - prev)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return state
    delta = state - prev
    dist.all_reduce(delta, op=dist.ReduceOp.SUM)
    state.copy_(prev + delta)
    return state
