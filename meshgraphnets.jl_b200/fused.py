"""build_graph (src/graph.jl:75-97) and the output post-processing of ode_step (src/solve.jl:205-218) handed to the
library as RECIPES (mgn_fused_io of include/mgn_b200.h) instead of being materialised: the encoders normalise and
concatenate while they stage their first operand, the decoder de-normalises and masks in its last epilogue, and the
pullback applies the transposed maps.  One model launch sequence replaces ~15 small launches per
right-hand-side evaluation."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .core import (Model, NormaliserOfflineMeanStd, NormaliserOfflineMinMax, NormaliserOnline, _dev_f32, _ptr, _stream,
                   call, graph_index_for)


def _seg(norm, x, col, width, inverse=False):
    s = _lib.FeatureSeg()
    s.d_x = x.data_ptr() if x is not None else None
    s.ld = int(x.shape[1]) if x is not None else 0
    s.col, s.width = int(col), int(width)
    if isinstance(norm, NormaliserOnline):
        if norm.dim != width:
            raise ValueError(f"NormaliserOnline({norm.dim}) applied to {width} features")
        s.kind, s.d_state, s.std_eps = _lib.FEAT_ONLINE, norm.state.data_ptr(), norm.std_epsilon
    elif isinstance(norm, (NormaliserOfflineMinMax, NormaliserOfflineMeanStd)) or norm is None:
        if norm is None:
            a, c = 1.0, 0.0
        elif isinstance(norm, NormaliserOfflineMinMax):
            a, c = norm._affine()
        else:
            a, c = float(1.0 / norm.std), float(-norm.mean / norm.std)
        if inverse:
            a, c = 1.0 / a, -c / a
        s.kind, s.scale, s.shift = _lib.FEAT_AFFINE, a, c
    else:
        raise TypeError(f"unsupported normaliser {type(norm).__name__}")
    return s


def update_online(jobs):
    """jobs: [(norm, x [rows, ld], col, width)] - the accumulate branch of every online normaliser of a step, two launches."""
    jobs = [(n, x, c, w) for n, x, c, w in jobs if isinstance(n, NormaliserOnline)]
    for i in range(0, len(jobs), 8):
        chunk = jobs[i:i + 8]
        arr = (_lib.NormUpdate * len(chunk))()
        for u, (n, x, c, w) in zip(arr, chunk):
            u.d_x, u.rows, u.ld, u.col, u.features = x.data_ptr(), int(x.shape[0]), int(x.shape[1]), int(c), int(w)
            u.d_state, u.max_acc = n.state.data_ptr(), n.max_acc
        call("mgn_norm_online_update_multi", arr, len(chunk), _stream())


class FusedGraph:
    """The recipe of one FeatureGraph: node blocks [(normaliser, matrix, col, width)], edge block, optional output blocks
    [(normaliser, width)] and val_mask.  Keeps the tensors it points into alive."""

    def __init__(self, node_blocks, edge_block, senders, receivers, n_nodes, out_blocks=(), val_mask=None, index_base=1):
        self.node_blocks = [(n, _dev_f32(x, "node block"), int(c), int(w)) for n, x, c, w in node_blocks]
        n, x, c, w = edge_block
        self.edge_block = (n, _dev_f32(x, "edge block"), int(c), int(w))
        self.out_blocks = list(out_blocks)
        self.val_mask = _dev_f32(val_mask, "val_mask") if val_mask is not None else None
        self.senders, self.receivers, self.n_nodes, self.index_base = senders, receivers, int(n_nodes), index_base
        io = _lib.FusedIo()
        io.n_node_segs = len(self.node_blocks)
        for i, (nm, xx, cc, ww) in enumerate(self.node_blocks):
            io.node[i] = _seg(nm, xx, cc, ww)
        io.n_edge_segs = 1
        io.edge[0] = _seg(*self.edge_block)
        io.n_out_segs = len(self.out_blocks)
        col = 0
        for i, (nm, ww) in enumerate(self.out_blocks):
            io.out[i] = _seg(nm, None, col, ww, inverse=True)
            col += ww
        io.d_val_mask = self.val_mask.data_ptr() if self.val_mask is not None else None
        self.io = io

    @property
    def index(self):
        return graph_index_for(self.n_nodes, self.senders, self.receivers, self.index_base)

    def accumulate(self):
        """What calling the normalisers does in build_graph when they still accumulate (src/graph.jl:80,84,93)."""
        update_online(self.node_blocks + [self.edge_block])


def forward_fused(model: Model, fg: FusedGraph, ps, training=False, slot=0):
    gi = fg.index
    ws = model.workspace(gi, training, slot)
    out = torch.empty((fg.n_nodes, model.cfg.out_dim), dtype=torch.float32, device=ps.device)
    call("mgn_forward_fused", model._h, gi._h, _ptr(ps), C.byref(fg.io), _ptr(out), _ptr(ws), ws.numel(), int(training),
         _stream())
    return out


def backward_fused(model: Model, fg: FusedGraph, ps, dout, want_dx=False, slot=0):
    gi = fg.index
    ws = model.workspace(gi, True, slot)
    dps = torch.empty(model.n_params, dtype=torch.float32, device=ps.device)
    dx = torch.empty((fg.n_nodes, model.cfg.node_in), dtype=torch.float32, device=ps.device) if want_dx else None
    call("mgn_backward_fused", model._h, gi._h, _ptr(ps), C.byref(fg.io), _ptr(_dev_f32(dout, "dout")), _ptr(dps),
         _ptr(dx), _ptr(ws), ws.numel(), _stream())
    return dps, dx
