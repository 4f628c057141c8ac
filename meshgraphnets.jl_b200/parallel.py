"""Data-parallel plumbing of the training step: one process per GPU, every rank
owns a strided subset of the (shuffled) time windows of a trajectory (src/MeshGraphNets.jl:364-370
iterates them sequentially; P-way DP is the `batchsize` the reference leaves unimplemented, :224), the
flat Float32 gradient is all-reduced and averaged, and the online-normaliser statistics are summed.
torch.distributed is only the transport (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


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


def pack_normaliser_states(norms):
    """Moves the device state of several NormaliserOnline objects into ONE flat buffer (each `.state` becomes a view of
    it), so that their data-parallel merge is a single all-reduce."""
    flat = torch.cat([n.state.reshape(-1) for n in norms])
    off = 0
    for n in norms:
        k = n.state.numel()
        n.state = flat[off:off + k]
        off += k
    return flat


class Communicator:
    """mgn_comm handle: the library's own NCCL transport (mgn_dp_* of include/mgn_b200.h) - what a Julia caller has,
    since it cannot use torch.distributed.  The 128-byte NCCL id is created by rank 0 (mgn_dp_unique_id) and handed to
    the other ranks by the caller; `from_torch_distributed` uses an existing process group for that one broadcast."""

    def __init__(self, unique_id: bytes, rank: int, world: int):
        if len(unique_id) != _lib.DP_UNIQUE_ID_BYTES:
            raise ValueError("the NCCL unique id is 128 bytes")
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        _lib.call("mgn_dp_init", C.c_char_p(unique_id), self.rank, self.world, C.byref(h))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.DP_UNIQUE_ID_BYTES)
        _lib.call("mgn_dp_unique_id", buf)
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device=None):
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, device=device)
        return cls(box[0], rank, world)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def allreduce_(self, t: torch.Tensor, op=_lib.DP_SUM):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise TypeError("allreduce_ needs a contiguous CUDA float32 tensor")
        _lib.call("mgn_dp_allreduce", self._h, C.c_void_p(t.data_ptr()), t.numel(), int(op), self._stream())
        return t

    def allreduce_sum_(self, t):
        return self.allreduce_(t, _lib.DP_SUM)

    def allreduce_mean_(self, t):
        return self.allreduce_(t, _lib.DP_MEAN)

    def allreduce_normaliser_(self, state: torch.Tensor, prev: torch.Tensor):
        """state = prev + sum_r (state_r - prev) on the device (see allreduce_normaliser_ above)."""
        _lib.call("mgn_dp_allreduce_normaliser", self._h, C.c_void_p(state.data_ptr()), C.c_void_p(prev.data_ptr()),
                  state.numel(), self._stream())
        return state

    def halo_exchange(self, send: torch.Tensor, send_rows, recv: torch.Tensor, recv_rows, row_bytes: int):
        n = self.world
        a = (C.c_int64 * n)(*[int(v) for v in send_rows])
        b = (C.c_int64 * n)(*[int(v) for v in recv_rows])
        _lib.call("mgn_halo_exchange", self._h, C.c_void_p(send.data_ptr()), a, C.c_void_p(recv.data_ptr()), b,
                  int(row_bytes), self._stream())
        return recv

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().mgn_dp_finalize(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
