"""Graph partitioning with halo exchange for meshes too large for one GPU (BASELINE
config 5).  The reference is single device (src/MeshGraphNets.jl:257) - this is new.

A rank owns a contiguous block of node ids (order the nodes locality-preserving first: grid slabs,
space-filling curve) and EVERY edge whose receiver it owns, so the scatter-sum of the messages stays
rank-local and deterministic.  Sender nodes owned elsewhere are appended to the local node list as halo
rows.  Between message-passing stages the owners' fresh latent rows overwrite the halo rows of their
readers (forward) and the halo rows of the latent gradient are returned to their owners and added
(backward).  The transport is pluggable: torch.distributed all-to-all (NCCL over NVLink on GPUs, gloo
in CPU tests) or an in-process exchange between P logical ranks on one device (tests)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .core import FeatureGraph, Model


@dataclass
class LocalGraph:
    rank: int
    lo: int                      # owned global nodes [lo, hi)
    hi: int
    halo_global: np.ndarray      # global ids (0-based) of the halo rows, ascending
    senders: np.ndarray          # local ids, 1-based Int32 (owned first, then halo)
    receivers: np.ndarray
    edge_ids: np.ndarray         # original edge ids of the local edges, ascending
    send_rows: dict = field(default_factory=dict)   # peer -> local (owned) row ids it needs, 0-based
    recv_rows: dict = field(default_factory=dict)   # peer -> local halo row ids it fills, same order as the peer's send

    @property
    def n_own(self):
        return self.hi - self.lo

    @property
    def n_local(self):
        return self.n_own + len(self.halo_global)

    def local_nodes_global(self):
        return np.concatenate([np.arange(self.lo, self.hi), self.halo_global])


def partition_bounds(n_nodes, world):
    return [(n_nodes * r) // world for r in range(world + 1)]


def build_partition(n_nodes, senders, receivers, world, index_base=1):
    """-> [LocalGraph] for ranks 0..world-1.  Integer work only (numpy, exact)."""
    s = np.asarray(senders, np.int64) - index_base
    r = np.asarray(receivers, np.int64) - index_base
    b = partition_bounds(n_nodes, world)
    owner = np.searchsorted(np.asarray(b[1:]), np.arange(n_nodes), side="right")
    parts = []
    for k in range(world):
        lo, hi = b[k], b[k + 1]
        eid = np.nonzero((r >= lo) & (r < hi))[0]
        ls, lr = s[eid], r[eid]
        halo = np.unique(ls[(ls < lo) | (ls >= hi)])
        loc = np.full(n_nodes, -1, np.int64)
        loc[lo:hi] = np.arange(hi - lo)
        loc[halo] = (hi - lo) + np.arange(len(halo))
        g = LocalGraph(k, lo, hi, halo, (loc[ls] + 1).astype(np.int32), (loc[lr] + 1).astype(np.int32), eid)
        for p in range(world):
            if p == k:
                continue
            mine = halo[owner[halo] == p]                      # my halo rows owned by p, ascending global id
            if len(mine):
                g.recv_rows[p] = loc[mine].astype(np.int32)
        parts.append(g)
    for k in range(world):                                      # the matching send lists
        for p, rows in parts[k].recv_rows.items():
            glob = parts[k].halo_global[rows - parts[k].n_own]
            parts[p].send_rows[k] = (glob - parts[p].lo).astype(np.int32)
    return parts


def build_partition_rank(n_nodes, senders, receivers, world, rank, index_base=1):
    """The LocalGraph of ONE rank (what a process under torchrun needs), without building the others: the send
    lists follow from the edges whose sender this rank owns and whose receiver a peer owns."""
    s = np.asarray(senders, np.int64) - index_base
    r = np.asarray(receivers, np.int64) - index_base
    b = np.asarray(partition_bounds(n_nodes, world))
    lo, hi = int(b[rank]), int(b[rank + 1])
    own_r = (r >= lo) & (r < hi)
    eid = np.nonzero(own_r)[0]
    ls, lr = s[eid], r[eid]
    halo = np.unique(ls[(ls < lo) | (ls >= hi)])
    halo_owner = np.searchsorted(b[1:], halo, side="right")
    g = LocalGraph(rank, lo, hi, halo, None, None, eid)
    pos = np.searchsorted(halo, ls)                              # local id of every sender
    is_halo = (ls < lo) | (ls >= hi)
    g.senders = (np.where(is_halo, (hi - lo) + pos, ls - lo) + 1).astype(np.int32)
    g.receivers = (lr - lo + 1).astype(np.int32)
    for p in range(world):
        if p == rank:
            continue
        rows = np.nonzero(halo_owner == p)[0]
        if len(rows):
            g.recv_rows[p] = ((hi - lo) + rows).astype(np.int32)
    mine_s = (s >= lo) & (s < hi) & ~own_r                       # my nodes read by edges that live elsewhere
    dst_owner = np.searchsorted(b[1:], r[mine_s], side="right")
    src = s[mine_s]
    for p in range(world):
        if p == rank:
            continue
        need = np.unique(src[dst_owner == p])
        if len(need):
            g.send_rows[p] = (need - lo).astype(np.int32)
    return g


class LocalExchange:
    """Transport between P logical ranks living in one process (tests, single-GPU dry runs)."""

    def __call__(self, sends, direction):
        """sends: {rank: {peer: tensor}} -> {rank: {peer: tensor received from that peer}}."""
        out = {}
        for r, d in sends.items():
            for p, t in d.items():
                out.setdefault(p, {})[r] = t
        return out


class DistExchange:
    """Transport over torch.distributed (one process per GPU): a single all_to_all_single per exchange with
    static split sizes known to both sides from the partition plan (NCCL over NVLink; gloo in CPU tests)."""

    def __init__(self, part: LocalGraph, world, device, model: Model = None):
        """`model` fixes the row widths (halo_row_bytes of the latent / of the gradient) for ranks that only receive in
        one direction (possible with one-way edge lists); without it the width is read off the first send buffer."""
        self.rank, self.world, self.device = part.rank, world, device
        self.width = None if model is None else {"fwd": model.halo_row_bytes(_lib.HALO_LATENT),
                                                 "bwd": model.halo_row_bytes(_lib.HALO_GRAD)}
        self.n_send = [len(part.send_rows.get(p, ())) for p in range(world)]   # forward: owned rows -> peers
        self.n_recv = [len(part.recv_rows.get(p, ())) for p in range(world)]   # forward: peers' rows -> my halo

    def __call__(self, sends, direction):
        import torch.distributed as dist
        mine = sends.get(self.rank, {})
        n_in, n_out = (self.n_send, self.n_recv) if direction == "fwd" else (self.n_recv, self.n_send)
        if self.width is not None:
            width = self.width[direction]
        elif mine:
            width = next(iter(mine.values())).shape[1]
        elif sum(n_out) == 0:
            width = 1                        # nothing moves either way on this rank
        else:
            raise _lib.MgnError(-1, "DistExchange: this rank sends nothing in this direction, so the row width is "
                                    "unknown - construct it with model=...")
        chunks = [mine[p] for p in range(self.world) if n_in[p]]
        inp = torch.cat(chunks) if chunks else torch.empty((0, width), dtype=torch.uint8, device=self.device)
        outp = torch.empty((sum(n_out), width), dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(outp, inp, list(n_out), list(n_in))
        recv, off = {}, 0
        for p in range(self.world):
            if n_out[p]:
                recv[p] = outp[off:off + n_out[p]]
                off += n_out[p]
        return {self.rank: recv}


class AbiExchange:
    """The same exchange through the library's own NCCL transport (mgn_halo_exchange: grouped ncclSend / ncclRecv on the
    caller's stream) - the path a Julia caller has.  `comm` is a parallel.Communicator."""

    def __init__(self, part: LocalGraph, world, comm, model: Model):
        self.rank, self.world, self.comm = part.rank, world, comm
        self.width = {"fwd": model.halo_row_bytes(_lib.HALO_LATENT), "bwd": model.halo_row_bytes(_lib.HALO_GRAD)}
        self.n_send = [len(part.send_rows.get(p, ())) for p in range(world)]
        self.n_recv = [len(part.recv_rows.get(p, ())) for p in range(world)]

    def __call__(self, sends, direction):
        mine = sends.get(self.rank, {})
        n_in, n_out = (self.n_send, self.n_recv) if direction == "fwd" else (self.n_recv, self.n_send)
        width = self.width[direction]
        dev = next(iter(mine.values())).device if mine else torch.device("cuda", torch.cuda.current_device())
        chunks = [mine[p] for p in range(self.world) if n_in[p]]
        inp = torch.cat(chunks) if chunks else torch.empty((0, width), dtype=torch.uint8, device=dev)
        outp = torch.empty((sum(n_out), width), dtype=torch.uint8, device=dev)
        self.comm.halo_exchange(inp, n_in, outp, n_out, width)
        recv, off = {}, 0
        for p in range(self.world):
            if n_out[p]:
                recv[p] = outp[off:off + n_out[p]]
                off += n_out[p]
        return {self.rank: recv}


class PartitionedModel:
    """One rank's share of a partitioned mesh: the local FeatureGraph, the exchange plan on the device and the
    stage-wise forward / backward of include/mgn_b200.h."""

    def __init__(self, model: Model, part: LocalGraph, nf_local, ef_local, device="cuda"):
        self.model, self.part = model, part
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.graph = FeatureGraph(nf_local, ef_local, dev(part.senders), dev(part.receivers))
        self.send = {p: dev(v) for p, v in part.send_rows.items()}
        self.recv = {p: dev(v) for p, v in part.recv_rows.items()}
        self.device = device

    def _buf(self, n, what):
        return torch.empty((n, self.model.halo_row_bytes(what)), dtype=torch.uint8, device=self.device)

    # forward: owners pack the latent read by MP step `step`; readers unpack into their halo rows
    def pack_latent(self, step, training):
        out = {}
        for p, rows in self.send.items():
            out[p] = self._buf(rows.numel(), _lib.HALO_LATENT)
            self.model.halo_rows(self.graph, training, _lib.HALO_LATENT, step, rows, out[p], _lib.ROWS_PACK)
        return out

    def unpack_latent(self, step, training, recv):
        for p, buf in recv.items():
            self.model.halo_rows(self.graph, training, _lib.HALO_LATENT, step, self.recv[p], buf, _lib.ROWS_UNPACK)

    # backward: readers pack (and zero) the halo rows of the latent gradient; owners add them
    def pack_grad(self):
        out = {}
        for p, rows in self.recv.items():
            out[p] = self._buf(rows.numel(), _lib.HALO_GRAD)
            self.model.halo_rows(self.graph, True, _lib.HALO_GRAD, 0, rows, out[p], _lib.ROWS_PACK_ZERO)
        return out

    def add_grad(self, recv):
        for p, buf in recv.items():
            self.model.halo_rows(self.graph, True, _lib.HALO_GRAD, 0, self.send[p], buf, _lib.ROWS_ADD)


def run_partitioned_step(ranks, ps, targets, masks, n_mask_total, exchange, loss_fn):
    """Lock-step training step over a list of PartitionedModel (all logical ranks of one process; with
    torch.distributed the list has one entry and `exchange` talks to the peers).
    targets / masks: per rank, local (owned rows first); masks hold 1-based local ids of OWNED nodes.
    Returns (per-rank partial parameter gradients, per-rank loss contributions, per-rank outputs)."""
    mps = ranks[0].model.cfg.mps
    E, D = _lib.STAGE_ENCODE, _lib.STAGE_DECODE
    outs = []
    for pm in ranks:
        pm.model.forward_stage(pm.graph, ps, E, training=True)
    for k in range(mps):
        for pm in ranks:
            pm.model.forward_stage(pm.graph, ps, k, training=True)
        if k + 1 < mps:   # the decoder only reads owned rows
            recv = exchange({pm.part.rank: pm.pack_latent(k + 1, True) for pm in ranks}, "fwd")
            for pm in ranks:
                pm.unpack_latent(k + 1, True, recv.get(pm.part.rank, {}))
    for pm in ranks:
        outs.append(pm.model.forward_stage(pm.graph, ps, D, training=True))
    grads, losses = [], []
    for pm, out, tgt, mask in zip(ranks, outs, targets, masks):
        loss, dout = loss_fn(out, tgt, mask, n_mask_total)
        dps = torch.zeros(pm.model.n_params, dtype=torch.float32, device=out.device)
        pm.model.backward_stage(pm.graph, ps, D, dps, dout=dout)
        grads.append(dps)
        losses.append(loss)
    for k in range(mps - 1, -1, -1):
        for pm, dps in zip(ranks, grads):
            pm.model.backward_stage(pm.graph, ps, k, dps)
        recv = exchange({pm.part.rank: pm.pack_grad() for pm in ranks}, "bwd")
        for pm in ranks:
            pm.add_grad(recv.get(pm.part.rank, {}))
    for pm, dps in zip(ranks, grads):
        pm.model.backward_stage(pm.graph, ps, E, dps)
    return grads, losses, outs


def masked_mse_partial(out, target, mask, n_mask_total, index_base=1):
    """This rank's share of the step! loss (src/strategies.jl:421) over its OWNED masked nodes, normalised by the
    global mask count: (sum over my masked rows) / n_mask_total, and the matching d(loss)/d(out)."""
    from .core import _ptr, _stream, call
    loss = torch.zeros(1, dtype=torch.float32, device=out.device)
    dout = torch.zeros_like(out)
    n = int(mask.numel())
    if n == 0:
        return loss, dout
    call("mgn_loss_mse_masked", _ptr(out), _ptr(target), out.shape[0], out.shape[1], _ptr(mask), n, int(index_base),
         _ptr(loss), _ptr(dout), _stream())
    scale = float(n) / float(n_mask_total)
    call("mgn_affine_apply", _ptr(dout), dout.shape[0], dout.shape[1], scale, 0.0, _ptr(dout), dout.shape[1], 0,
         _stream())
    call("mgn_affine_apply", _ptr(loss), 1, 1, scale, 0.0, _ptr(loss), 1, 0, _stream())
    return loss, dout
