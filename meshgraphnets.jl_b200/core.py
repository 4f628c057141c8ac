"""Host-side mirror of the GraphNetCore.jl names MeshGraphNets.jl uses (docs/src/graph_net_core.md:5-36),
written above the C ABI of libmgn_b200.so.  Julia is not in this image, so this Python layer plays
the role of julia/GraphNetCoreB200.jl: same names, argument meaning and error behaviour, so that
the parity tests read like the reference's call sites.  PyTorch is only plumbing here (device
memory, streams); all arithmetic runs in the library's CUDA kernels.

Layout note: Julia matrices are column-major ``(features, entities)``; tensors here are the same
bytes viewed row-major ``[entities, features]``.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import COMPUTE_BF16, COMPUTE_FP32, MgnError, ModelConfig, ParamEntry, call


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data_as(C.c_void_p)
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(x, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
        raise TypeError(f"{what} must be a CUDA float32 tensor (CuArray{{Float32}})")
    return x.contiguous()


def _dev_i32(x, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.int32):
        raise TypeError(f"{what} must be a CUDA int32 tensor (CuArray{{Int32}})")
    return x.contiguous()


# ------------------------------------------------------------------------------------------------
# Utilities: one_hot / triangles_to_edges / parse_edges  (src/graph.jl:26,30,38)
# ------------------------------------------------------------------------------------------------


def one_hot(v, depth, offset=0):
    """GraphNetCore.one_hot(v, depth, offset) as called at src/graph.jl:26-27; returns [n, depth]."""
    v = np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(-1))
    out = np.empty((v.shape[0], int(depth)), dtype=np.float32)
    call("mgn_one_hot", _ptr(v), v.shape[0], int(depth), int(offset), _ptr(out))
    return out


def triangles_to_edges(cells):
    """GraphNetCore.triangles_to_edges(cells) as called at src/graph.jl:30; cells is [C, 3]."""
    cells = np.ascontiguousarray(np.asarray(cells, dtype=np.int32).reshape(-1, 3))
    n = cells.shape[0]
    s = np.empty(6 * n, dtype=np.int32)
    r = np.empty(6 * n, dtype=np.int32)
    ne = C.c_int64(0)
    call("mgn_triangles_to_edges", _ptr(cells), n, _ptr(s), _ptr(r), C.byref(ne))
    return s[:ne.value].copy(), r[:ne.value].copy()


def parse_edges(edges):
    """GraphNetCore.parse_edges(edges) as called at src/graph.jl:38; edges is [U, 2]."""
    edges = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
    n = edges.shape[0]
    s = np.empty(2 * n, dtype=np.int32)
    r = np.empty(2 * n, dtype=np.int32)
    call("mgn_parse_edges", _ptr(edges), n, _ptr(s), _ptr(r))
    return s, r


def shift_one_based(senders, receivers):
    """The in-place ``.+= 1`` of src/graph.jl:31-34; returns True when the shift was applied."""
    flag = C.c_int32(0)
    call("mgn_shift_one_based", _ptr(senders), _ptr(receivers), senders.shape[0], C.byref(flag))
    return bool(flag.value)


def edge_features(mesh_pos, senders, receivers, index_base=1):
    """[rel ; ||rel||] of src/graph.jl:35-36,49-52 (host); mesh_pos [N, dim] -> [E, dim+1]."""
    pos = np.ascontiguousarray(np.asarray(mesh_pos, dtype=np.float32))
    s = np.ascontiguousarray(senders, dtype=np.int32)
    r = np.ascontiguousarray(receivers, dtype=np.int32)
    out = np.empty((s.shape[0], pos.shape[1] + 1), dtype=np.float32)
    call("mgn_edge_features", _ptr(pos), pos.shape[0], pos.shape[1], _ptr(s), _ptr(r), s.shape[0],
         int(index_base), _ptr(out))
    return out


def mse_reduce(target, outputs):
    """GraphNetCore.mse_reduce - only its identity matters: step_ dispatches on it."""
    raise MgnError(-1, "mse_reduce is a marker for step_(); the loss runs inside libmgn_b200")


# ------------------------------------------------------------------------------------------------
# Graph handle + FeatureGraph (src/graph.jl:87-96)
# ------------------------------------------------------------------------------------------------


class GraphIndex:
    """Device CSR/CSC of a (senders, receivers) pair - mgn_graph handle."""

    def __init__(self, n_nodes, senders, receivers, index_base=1):
        senders = _dev_i32(senders, "senders")
        receivers = _dev_i32(receivers, "receivers")
        if senders.shape != receivers.shape:
            raise ValueError("senders and receivers differ in length")
        self.n_nodes = int(n_nodes)
        self.n_edges = int(senders.shape[0])
        self.index_base = index_base
        h = C.c_void_p()
        call("mgn_graph_create", self.n_nodes, self.n_edges, _ptr(senders), _ptr(receivers),
             int(index_base), _stream(), C.byref(h))
        self._h = h

    def index_arrays(self):
        rp = np.empty(self.n_nodes + 1, np.int32)
        cp = np.empty(self.n_nodes + 1, np.int32)
        perm = np.empty(self.n_edges, np.int32)
        perm_s = np.empty(self.n_edges, np.int32)
        call("mgn_graph_get_index", self._h, _ptr(rp), _ptr(perm), _ptr(cp), _ptr(perm_s))
        return rp, perm, cp, perm_s

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().mgn_graph_destroy(h)
            except Exception:
                pass


_graph_cache: dict = {}


def graph_index_for(n_nodes, senders, receivers, index_base=1):
    """FeatureGraphs of one trajectory share senders/receivers (src/MeshGraphNets.jl:360-370), so the
    CSR is built once per (buffer, size) and reused by every build_graph call."""
    key = (senders.data_ptr(), receivers.data_ptr(), int(n_nodes), int(senders.shape[0]), index_base)
    hit = _graph_cache.get(key)
    if hit is None:
        if len(_graph_cache) > 16:
            _graph_cache.clear()
        # the entry keeps the index tensors alive, so their addresses cannot be recycled for a
        # different edge list while the entry exists
        hit = (GraphIndex(n_nodes, senders, receivers, index_base), senders, receivers)
        _graph_cache[key] = hit
    return hit[0]


@dataclass
class FeatureGraph:
    """GraphNetCore.FeatureGraph(nf, ef, senders, receivers) as built at src/graph.jl:87-96."""
    node_features: torch.Tensor   # [N, node_in]
    edge_features: torch.Tensor   # [E, edge_in]
    senders: torch.Tensor         # Int32 [E]
    receivers: torch.Tensor       # Int32 [E]
    index_base: int = 1

    @property
    def index(self):
        # memoised on the object: the handle (and the workspaces keyed by it) must stay the same between the
        # forward and backward stages of a step even if the global cache is recycled by other graphs in between
        key = (self.senders.data_ptr(), self.receivers.data_ptr(), int(self.node_features.shape[0]),
               int(self.senders.shape[0]), self.index_base)
        hit = self.__dict__.get("_gi")
        if hit is None or hit[0] != key:
            hit = (key, graph_index_for(self.node_features.shape[0], self.senders, self.receivers, self.index_base))
            self.__dict__["_gi"] = hit
        return hit[1]


# ------------------------------------------------------------------------------------------------
# Normalisers (constructed at src/MeshGraphNets.jl:74-206)
# ------------------------------------------------------------------------------------------------


class NormaliserOfflineMinMax:
    """GraphNetCore.NormaliserOfflineMinMax(data_min, data_max[, target_min, target_max])."""

    def __init__(self, data_min, data_max, target_min=0.0, target_max=1.0):
        f = np.float32
        self.data_min, self.data_max = f(data_min), f(data_max)
        self.target_min, self.target_max = f(target_min), f(target_max)

    def _affine(self):
        a = (self.target_max - self.target_min) / (self.data_max - self.data_min)
        return float(a), float(self.target_min - self.data_min * a)

    def __call__(self, x, out=None, col=0):
        a, c = self._affine()
        return _affine(x, a, c, out, col)

    def inverse(self, y, out=None, col=0):
        a, c = self._affine()
        return _affine(y, 1.0 / a, -c / a, out, col)

    def apply_ld(self, x, col_x, width, mode, out, col_out):
        """Strided map / transposed Jacobian (``_lib.NORM_*`` modes) used by the solver strategies."""
        a, c = self._affine()
        return _affine_ld(x, col_x, width, *_affine_mode(a, c, mode), out, col_out)


class NormaliserOfflineMeanStd:
    """GraphNetCore.NormaliserOfflineMeanStd(mean, std)."""

    def __init__(self, mean, std):
        self.mean, self.std = np.float32(mean), np.float32(std)

    def __call__(self, x, out=None, col=0):
        return _affine(x, float(1.0 / self.std), float(-self.mean / self.std), out, col)

    def inverse(self, y, out=None, col=0):
        return _affine(y, float(self.std), float(self.mean), out, col)

    def apply_ld(self, x, col_x, width, mode, out, col_out):
        a, c = float(1.0 / self.std), float(-self.mean / self.std)
        return _affine_ld(x, col_x, width, *_affine_mode(a, c, mode), out, col_out)


def _affine(x, a, c, out, col):
    x = _dev_f32(x, "x")
    rows, F = x.shape
    if out is None:
        out, col = torch.empty_like(x), 0
    call("mgn_affine_apply", _ptr(x), rows, F, a, c, _ptr(out), out.shape[1], col, _stream())
    return out


def _affine_mode(a, c, mode):
    """(scale, shift) of y = a x + c in the four ``_lib.NORM_*`` modes."""
    return {_lib.NORM_FORWARD: (a, c), _lib.NORM_INVERSE: (1.0 / a, -c / a),
            _lib.NORM_FORWARD_VJP: (a, 0.0), _lib.NORM_INVERSE_VJP: (1.0 / a, 0.0)}[mode]


def _affine_ld(x, col_x, width, scale, shift, out, col_out):
    x, out = _dev_f32(x, "x"), _dev_f32(out, "out")
    call("mgn_affine_apply_ld", _ptr(x), x.shape[1], int(col_x), x.shape[0], int(width), float(scale), float(shift),
         _ptr(out), out.shape[1], int(col_out), _stream())
    return out


class NormaliserOnline:
    """GraphNetCore.NormaliserOnline(dim, device; max_acc, std_epsilon): accumulates on call
    (src/graph.jl:80,84,93; src/strategies.jl:399-410).  State lives on the device as
    [acc_sum | acc_sum_sq | acc_count | num_acc]."""

    def __init__(self, dim, device="cuda", max_acc=1.0e6, std_epsilon=1e-8):
        self.dim = int(dim)
        self.max_acc = float(max_acc)
        self.std_epsilon = float(std_epsilon)
        self.state = torch.zeros(2 * self.dim + 2, dtype=torch.float32, device=device)

    def __call__(self, x, accumulate=True, out=None, col=0):
        x = _dev_f32(x, "x")
        rows, F = x.shape
        if F != self.dim:
            raise ValueError(f"NormaliserOnline({self.dim}) applied to {F} features")
        if accumulate:
            call("mgn_norm_online_update", _ptr(x), rows, F, _ptr(self.state), self.max_acc, _stream())
        if out is None:
            out, col = torch.empty_like(x), 0
        call("mgn_norm_online_apply", _ptr(x), rows, F, _ptr(self.state), self.std_epsilon, 0, _ptr(out),
             out.shape[1], col, _stream())
        return out

    def inverse(self, y, out=None, col=0):
        y = _dev_f32(y, "y")
        rows, F = y.shape
        if out is None:
            out, col = torch.empty_like(y), 0
        call("mgn_norm_online_apply", _ptr(y), rows, F, _ptr(self.state), self.std_epsilon, 1, _ptr(out),
             out.shape[1], col, _stream())
        return out


def _online_apply_ld(self, x, col_x, width, mode, out, col_out):
    """Strided map / transposed Jacobian from the CURRENT statistics; never accumulates."""
    x, out = _dev_f32(x, "x"), _dev_f32(out, "out")
    if width != self.dim:
        raise ValueError(f"NormaliserOnline({self.dim}) applied to {width} features")
    call("mgn_norm_online_apply_ld", _ptr(x), x.shape[1], int(col_x), x.shape[0], int(width), _ptr(self.state),
         self.std_epsilon, int(mode), _ptr(out), out.shape[1], int(col_out), _stream())
    return out


NormaliserOnline.apply_ld = _online_apply_ld


def inverse_data(norm, y):
    """GraphNetCore.inverse_data(norm, data) as called at src/solve.jl:207-209."""
    return norm.inverse(y)


# ------------------------------------------------------------------------------------------------
# Model
# ------------------------------------------------------------------------------------------------


class Model:
    """The Lux model object ``mgn.model``: callable as ``model(graph, ps, st) -> (out, st)``
    (src/solve.jl:200).  Holds the mgn_model handle and a per-graph workspace."""

    def __init__(self, node_in, edge_in, out_dim, mps, layer_size, hidden_layers, ln_eps=1e-5,
                 compute_mode=COMPUTE_FP32, dense_layers=0, ln_scale_first=False, aggregate_post_residual=False):
        """The last three arguments are the recalled GraphNetCore internals exposed as switches (SURVEY.md section 9,
        mgn_model_config): Dense layers per MLP (0 = hidden_layers + 2), LayerNorm parameter order, aggregation order."""
        self.cfg = ModelConfig(int(node_in), int(edge_in), int(out_dim), int(layer_size), int(mps),
                               int(hidden_layers), float(ln_eps), int(compute_mode), int(dense_layers),
                               int(bool(ln_scale_first)), int(bool(aggregate_post_residual)))
        h = C.c_void_p()
        call("mgn_model_create", C.byref(self.cfg), C.byref(h))
        self._h = h
        n = C.c_int64(0)
        call("mgn_model_param_count", h, C.byref(n))
        self.n_params = n.value
        self._ws = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().mgn_model_destroy(h)
            except Exception:
                pass

    def param_layout(self):
        n = C.c_int32(0)
        call("mgn_model_param_layout", self._h, None, 0, C.byref(n))
        arr = (ParamEntry * n.value)()
        call("mgn_model_param_layout", self._h, arr, n.value, C.byref(n))
        return [(e.name.decode(), e.offset, e.rows, e.cols) for e in arr]

    def workspace(self, gi: GraphIndex, training: bool, slot=0):
        """``slot`` > 0 gives further workspaces for the same graph: the stages of one Runge-Kutta step keep their
        activations side by side until the reverse sweep has consumed them (solver strategies)."""
        key = (id(gi), bool(training), int(slot))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = C.c_size_t(0)
            call("mgn_workspace_bytes", self._h, gi._h, int(training), C.byref(nbytes))
            if len(self._ws) > 16:   # evict other graphs only: the slots of this one may hold a step in flight
                for k in [k for k in self._ws if k[0] != id(gi)]:
                    del self._ws[k]
            ws = (torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device="cuda"), gi)
            self._ws[key] = ws
        return ws[0]

    def workspace_bytes(self, gi: GraphIndex, training: bool):
        nbytes = C.c_size_t(0)
        call("mgn_workspace_bytes", self._h, gi._h, int(training), C.byref(nbytes))
        return nbytes.value

    def forward(self, graph: FeatureGraph, ps, training=False, slot=0):
        gi = graph.index
        nf = _dev_f32(graph.node_features, "node_features")
        ef = _dev_f32(graph.edge_features, "edge_features")
        ps = _dev_f32(ps, "ps")
        if ps.numel() != self.n_params:
            raise ValueError(f"ps has {ps.numel()} elements, model needs {self.n_params}")
        if nf.shape[1] != self.cfg.node_in or ef.shape[1] != self.cfg.edge_in:
            raise ValueError("feature widths do not match the model")
        ws = self.workspace(gi, training, slot)
        out = torch.empty((nf.shape[0], self.cfg.out_dim), dtype=torch.float32, device=nf.device)
        call("mgn_forward", self._h, gi._h, _ptr(ps), _ptr(nf), _ptr(ef), _ptr(out), _ptr(ws), ws.numel(),
             int(training), _stream())
        return out

    def backward(self, graph: FeatureGraph, ps, dout, want_dnf=False, slot=0):
        """Pullback of the matching ``forward(..., training=True)``: (d_ps, d_nf or None)."""
        gi = graph.index
        ws = self.workspace(gi, True, slot)
        nf = _dev_f32(graph.node_features, "node_features")
        ef = _dev_f32(graph.edge_features, "edge_features")
        dout = _dev_f32(dout, "dout")
        dps = torch.empty(self.n_params, dtype=torch.float32, device=nf.device)
        dnf = torch.empty_like(nf) if want_dnf else None
        call("mgn_backward", self._h, gi._h, _ptr(ps), _ptr(nf), _ptr(ef), _ptr(dout), _ptr(dps),
             _ptr(dnf), _ptr(ws), ws.numel(), _stream())
        return dps, dnf

    def __call__(self, graph, ps, st=None):
        return self.forward(graph, ps, training=False), st

    # ---- stage-wise execution for graph-partitioned meshes (include/mgn_b200.h, "halo exchange") ----
    def forward_stage(self, graph: FeatureGraph, ps, stage, training=False, out=None):
        gi = graph.index
        ws = self.workspace(gi, training)
        nf = _dev_f32(graph.node_features, "node_features")
        ef = _dev_f32(graph.edge_features, "edge_features")
        if stage == _lib.STAGE_DECODE and out is None:
            out = torch.empty((nf.shape[0], self.cfg.out_dim), dtype=torch.float32, device=nf.device)
        call("mgn_forward_stage", self._h, gi._h, _ptr(ps), _ptr(nf), _ptr(ef), _ptr(out), _ptr(ws), ws.numel(),
             int(training), int(stage), _stream())
        return out

    def backward_stage(self, graph: FeatureGraph, ps, stage, dps, dout=None, dnf=None):
        gi = graph.index
        ws = self.workspace(gi, True)
        nf = _dev_f32(graph.node_features, "node_features")
        ef = _dev_f32(graph.edge_features, "edge_features")
        call("mgn_backward_stage", self._h, gi._h, _ptr(ps), _ptr(nf), _ptr(ef), _ptr(dout), _ptr(dps), _ptr(dnf),
             _ptr(ws), ws.numel(), int(stage), _stream())

    def halo_row_bytes(self, what):
        n = C.c_size_t(0)
        call("mgn_halo_row_bytes", self._h, int(what), C.byref(n))
        return n.value

    def halo_rows(self, graph: FeatureGraph, training, what, step, rows, buf, op):
        """rows: int32 CUDA tensor of 0-based local node ids; buf: contiguous CUDA byte buffer."""
        gi = graph.index
        ws = self.workspace(gi, training)
        call("mgn_halo_rows", self._h, gi._h, _ptr(ws), ws.numel(), int(training), int(what), int(step),
             _ptr(rows), int(rows.numel()), _ptr(buf), int(op), _stream())


def init_params(model: Model, seed=1234):
    """Lux default initialisation restated: glorot-uniform Dense weights, zero biases, LayerNorm
    scale 1 / bias 0.  Uses the same PCG64 stream as oracle.init_params so that both sides can be
    seeded identically in tests (Julia's own RNG stream is not reproduced; tensors are exchanged)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.zeros(model.n_params, dtype=np.float64)
    for name, off, rows, cols in model.param_layout():
        if name.endswith(".weight"):
            lim = math.sqrt(6.0 / (rows + cols))
            p[off:off + rows * cols] = rng.uniform(-lim, lim, size=rows * cols)
        elif name.endswith("layernorm.scale"):
            p[off:off + rows] = 1.0
    return torch.from_numpy(p.astype(np.float32))


class GraphNetwork:
    """GraphNetCore.GraphNetwork: fields .model .ps .st .e_norm .n_norm .o_norm
    (src/graph.jl:80-93, src/solve.jl:200-208, src/MeshGraphNets.jl:288,376-377)."""

    def __init__(self, model, ps, st, e_norm, n_norm, o_norm):
        self.model, self.ps, self.st = model, ps, st
        self.e_norm, self.n_norm, self.o_norm = e_norm, n_norm, o_norm


def build_model(quantities_size, dims, output_size, mps, layer_size, hidden_layers, device="cuda",
                compute_mode=COMPUTE_FP32, seed=1234):
    """GraphNetCore.build_model(quantities, dims, outputs, mps, layer_size, hidden_layers, device)
    (reached through ``load`` at src/MeshGraphNets.jl:282-285) -> (model, ps, st)."""
    model = Model(quantities_size, dims + 1, output_size, mps, layer_size, hidden_layers,
                  compute_mode=compute_mode)
    ps = init_params(model, seed).to(device)
    return model, ps, None


def step_(mgn: GraphNetwork, graph: FeatureGraph, target, mask, loss_function=mse_reduce, index_base=1):
    """GraphNetCore.step!(mgn, graph, target, mask, mse_reduce) -> (gs, loss) as called at
    src/strategies.jl:421: gs is a 1-tuple (iterated at src/MeshGraphNets.jl:375-377), loss a
    device scalar."""
    if loss_function is not mse_reduce:
        raise MgnError(-1, "only mse_reduce is supported as the step! loss")
    model: Model = mgn.model
    target = _dev_f32(target, "target")
    mask = _dev_i32(mask, "mask")
    out = model.forward(graph, mgn.ps, training=True)
    loss = torch.empty(1, dtype=torch.float32, device=out.device)
    dout = torch.empty_like(out)
    call("mgn_loss_mse_masked", _ptr(out), _ptr(target), out.shape[0], out.shape[1], _ptr(mask),
         mask.shape[0], int(index_base), _ptr(loss), _ptr(dout), _stream())
    gs, _ = model.backward(graph, mgn.ps, dout)
    return (gs,), loss


def step_dp_(mgn: GraphNetwork, graph: FeatureGraph, target, mask, opt=None, opt_state=None, comm=None, n_buckets=4,
             index_base=1):
    """step! (src/strategies.jl:421) fused with what the driver does with its result (src/MeshGraphNets.jl:374-378):
    forward, loss and mgn_backward_dp - the backward pass whose gradient buckets are all-reduced (mean over the ranks of
    `comm`, a parallel.Communicator) and fed to Adam on a side stream while the rest of the backward pass still runs.
    With `opt` the parameters are updated in place.  Returns ((gs,), loss) like step_."""
    model: Model = mgn.model
    target = _dev_f32(target, "target")
    mask = _dev_i32(mask, "mask")
    out = model.forward(graph, mgn.ps, training=True)
    loss = torch.empty(1, dtype=torch.float32, device=out.device)
    dout = torch.empty_like(out)
    call("mgn_loss_mse_masked", _ptr(out), _ptr(target), out.shape[0], out.shape[1], _ptr(mask),
         mask.shape[0], int(index_base), _ptr(loss), _ptr(dout), _stream())
    gi = graph.index
    ws = model.workspace(gi, True, 0)
    nf = _dev_f32(graph.node_features, "node_features")
    ef = _dev_f32(graph.edge_features, "edge_features")
    key = ("_dps", mgn.ps.data_ptr())
    dps = model.__dict__.get(key)
    if dps is None or dps.numel() != model.n_params:
        dps = torch.empty(model.n_params, dtype=torch.float32, device=nf.device)
        model.__dict__[key] = dps
    cfg = None
    if opt is not None:
        opt_state["t"] += 1
        cfg = _lib.AdamConfig(opt.eta, opt.beta[0], opt.beta[1], opt.epsilon, opt_state["m"].data_ptr(),
                              opt_state["v"].data_ptr(), opt_state["dev"].data_ptr())
    call("mgn_backward_dp", model._h, gi._h, _ptr(mgn.ps), _ptr(nf), _ptr(ef), _ptr(dout), _ptr(dps), None, _ptr(ws),
         ws.numel(), comm._h if comm is not None else None, C.byref(cfg) if cfg is not None else None, int(n_buckets),
         _stream())
    return (dps,), loss


class Adam:
    """Optimisers.Adam + Optimisers.setup/update as used at src/MeshGraphNets.jl:288,374-378."""

    def __init__(self, eta=1e-4, beta=(0.9, 0.999), epsilon=1e-8):
        self.eta, self.beta, self.epsilon = float(eta), (float(beta[0]), float(beta[1])), float(epsilon)

    def setup(self, ps):
        # "dev" = {int64 step; float c1; float c2}: the step counter lives on the device so that a
        # whole training step can be replayed as a CUDA graph
        return {"m": torch.zeros_like(ps), "v": torch.zeros_like(ps), "t": 0,
                "dev": torch.zeros(2, dtype=torch.int64, device=ps.device)}

    def update(self, state, ps, gs):
        """In place on ps / state (Optimisers.update returns new objects; same values)."""
        state["t"] += 1
        call("mgn_adam_step_device", _ptr(ps), _ptr(gs), _ptr(state["m"]), _ptr(state["v"]), ps.numel(),
             self.eta, self.beta[0], self.beta[1], self.epsilon, _ptr(state["dev"]), _stream())
        return state, ps


def profile_begin(tag=-1):
    """Starts counting the library's kernel launches (and timing the family `tag`)."""
    call("mgn_profile_begin", int(tag))


def profile_end():
    """-> (launches, tagged launches, tagged ms, {family name: launches})."""
    n, k, ms = C.c_int64(0), C.c_int64(0), C.c_float(0)
    per = (C.c_int64 * 32)()
    call("mgn_profile_end", C.byref(n), C.byref(k), C.byref(ms), per, 32)
    names = {}
    buf = C.create_string_buffer(64)
    for i in range(32):
        if _lib.load().mgn_profile_tag_name(i, buf, 64) != 0:
            break
        if per[i]:
            names[buf.value.decode()] = int(per[i])
    return n.value, k.value, ms.value, names


def profile_tag(name):
    buf = C.create_string_buffer(64)
    for i in range(32):
        if _lib.load().mgn_profile_tag_name(i, buf, 64) != 0:
            break
        if buf.value.decode() == name:
            return i
    raise KeyError(name)
