"""CPU oracle for the MeshGraphNets.jl Encode-Process-Decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``meshgraphnets.jl_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs do.  It is the checker, never the product path.

PARITY UNPINNED: the reference (una-auxme/MeshGraphNets.jl v0.4.1, /root/reference) holds
no golden vectors for this path (test/runtests.jl:10-18 is Aqua hygiene only) and the
arithmetic lives in the un-vendored dependency GraphNetCore.jl (Project.toml:11, compat
"0.3" Project.toml:36, no Manifest => version unpinned) which is absent from this machine,
as is a Julia toolchain.  This file therefore restates

  * the reference's own call sites (file:line cited per function), and
  * the published algorithm GraphNetCore.jl implements (Pfaff et al. 2021, "Learning
    mesh-based simulation with graph networks"; DeepMind meshgraphnets core_model.py /
    common.py / normalization.py) with Lux 0.5 Dense / LayerNorm semantics.

Every recalled semantic is a keyword knob with the recalled value as default (see
ASSUMPTIONS in DESIGN.md).  The integer path is pinned by hand-derived known-answer tests
in tests/test_oracle_kat.py.

All arrays use the C view of Julia's column-major ``(features, entities)`` matrices, i.e.
``[entities, features]`` row-major - byte-identical memory.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------------
# Integer path (bit-exact)
# ----------------------------------------------------------------------------------------------


def one_hot(v, depth, offset=0):
    """GraphNetCore.one_hot as called at src/graph.jl:26-27: row ``v+offset`` (1-based) of a
    ``depth x n`` Float32 matrix is 1.  Returned as [n, depth]; out-of-range rows stay zero."""
    v = np.asarray(v, dtype=np.int64).reshape(-1)
    out = np.zeros((v.shape[0], int(depth)), dtype=np.float32)
    for i, x in enumerate(v):
        j = int(x) + int(offset) - 1  # 1-based row -> 0-based column
        if 0 <= j < depth:
            out[i, j] = 1.0
    return out


def triangles_to_edges(cells):
    """GraphNetCore.triangles_to_edges as called at src/graph.jl:30 (DeepMind
    common.triangles_to_edges semantics): edges = [f0f1 for all faces; f1f2 ...; f2f0 ...],
    each stored as (max, min); unique keeping FIRST-OCCURRENCE order; two-way result
    senders=[max;min], receivers=[min;max].  ``cells`` is [C, 3] (C view of Julia's 3 x C)."""
    cells = np.asarray(cells, dtype=np.int32).reshape(-1, 3)
    raw = [(int(c[0]), int(c[1])) for c in cells]
    raw += [(int(c[1]), int(c[2])) for c in cells]
    raw += [(int(c[2]), int(c[0])) for c in cells]
    seen = set()
    hi, lo = [], []
    for a, b in raw:
        key = (max(a, b), min(a, b))
        if key in seen:
            continue
        seen.add(key)
        hi.append(key[0])
        lo.append(key[1])
    senders = np.asarray(hi + lo, dtype=np.int32)
    receivers = np.asarray(lo + hi, dtype=np.int32)
    return senders, receivers


def parse_edges(edges):
    """GraphNetCore.parse_edges as called at src/graph.jl:38: ``edges`` is [U, 2] (C view of
    2 x U); two-way result senders=[s;r], receivers=[r;s]."""
    edges = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
    s, r = edges[:, 0], edges[:, 1]
    return np.concatenate([s, r]).astype(np.int32), np.concatenate([r, s]).astype(np.int32)


def shift_to_one_based(senders, receivers):
    """src/graph.jl:31-34 (and :39-42): if 0 occurs anywhere, both arrays are incremented."""
    senders = np.array(senders, dtype=np.int32)
    receivers = np.array(receivers, dtype=np.int32)
    if (senders == 0).any() or (receivers == 0).any():
        senders += 1
        receivers += 1
    return senders, receivers


def create_edges_1d(n):
    """src/dataset.jl:379-382: 1-D chain, 1-based node pairs [i, i+1]."""
    i = np.arange(1, n, dtype=np.int32)
    return np.stack([i, i + 1], axis=1)


def edge_features(mesh_pos, senders, receivers, index_base=1):
    """src/graph.jl:35-36,49-52: rel = pos[:, s] - pos[:, r]; features = [rel ; ||rel||_2].
    Float32 subtraction; the norm follows LinearAlgebra.norm on a Float32 vector (generic
    norm2: accumulates in the widened type, rounds once)."""
    pos = np.asarray(mesh_pos, dtype=np.float32)
    s = np.asarray(senders, dtype=np.int64) - index_base
    r = np.asarray(receivers, dtype=np.int64) - index_base
    rel = (pos[s] - pos[r]).astype(np.float32)
    nrm = np.sqrt((rel.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    return np.concatenate([rel, nrm[:, None]], axis=1).astype(np.float32)


def node_mask(node_type, types_updated, index_base=1):
    """src/MeshGraphNets.jl:352: Int32 ids (1-based) of nodes whose type is in types_updated."""
    nt = np.asarray(node_type).reshape(-1)
    return (np.nonzero(np.isin(nt, list(types_updated)))[0] + index_base).astype(np.int32)


def val_mask(node_type, types_updated, out_dim):
    """src/MeshGraphNets.jl:354-358: Float32 0/1 mask repeated over the output rows -> [N, out]."""
    nt = np.asarray(node_type).reshape(-1)
    m = np.isin(nt, list(types_updated)).astype(np.float32)
    return np.repeat(m[:, None], out_dim, axis=1)


def build_csr(keys, n_nodes, index_base=1):
    """NEW (no reference counterpart): stable sort of edge ids by ``keys``
    (receivers for CSR, senders for CSC).  Stability makes the order inside a segment the
    ascending original edge id, i.e. the summation order of NNlib's sequential CPU scatter
    that GraphNetCore's aggregation uses.  Returns (row_ptr[N+1], perm[E]) with perm holding
    0-based original edge ids."""
    k = np.asarray(keys, dtype=np.int64) - index_base
    perm = np.argsort(k, kind="stable").astype(np.int32)
    counts = np.bincount(k, minlength=n_nodes)
    row_ptr = np.zeros(n_nodes + 1, dtype=np.int32)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr, perm


# ----------------------------------------------------------------------------------------------
# Normalisers (constructed src/MeshGraphNets.jl:74-206)
# ----------------------------------------------------------------------------------------------


class NormaliserOfflineMinMax:
    """GraphNetCore.NormaliserOfflineMinMax(data_min, data_max[, target_min, target_max]);
    built at src/MeshGraphNets.jl:81,102,117-134; default target range [0, 1]."""

    def __init__(self, data_min, data_max, target_min=0.0, target_max=1.0):
        self.data_min = np.float32(data_min)
        self.data_max = np.float32(data_max)
        self.target_min = np.float32(target_min)
        self.target_max = np.float32(target_max)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float32)
        return ((x - self.data_min) / (self.data_max - self.data_min) *
                (self.target_max - self.target_min) + self.target_min).astype(np.float32)

    def inverse(self, y):
        y = np.asarray(y, dtype=np.float32)
        return ((y - self.target_min) / (self.target_max - self.target_min) *
                (self.data_max - self.data_min) + self.data_min).astype(np.float32)

    def scale_shift(self, F):
        """Affine form y = x*a + c per feature (used by the fused build_graph kernel)."""
        a = (self.target_max - self.target_min) / (self.data_max - self.data_min)
        c = self.target_min - self.data_min * a
        return np.full(F, a, np.float32), np.full(F, c, np.float32)


class NormaliserOfflineMeanStd:
    """GraphNetCore.NormaliserOfflineMeanStd(mean, std); built at src/MeshGraphNets.jl:84,167."""

    def __init__(self, mean, std):
        self.mean = np.float32(mean)
        self.std = np.float32(std)

    def __call__(self, x):
        return ((np.asarray(x, np.float32) - self.mean) / self.std).astype(np.float32)

    def inverse(self, y):
        return (np.asarray(y, np.float32) * self.std + self.mean).astype(np.float32)


class NormaliserOnline:
    """GraphNetCore.NormaliserOnline(dim, device; max_acc, std_epsilon) - DeepMind
    normalization.py semantics; built at src/MeshGraphNets.jl:90,138,155,183; applied at
    src/graph.jl:80,84,93 and src/strategies.jl:399-410 (STATEFUL: accumulates on call)."""

    def __init__(self, dim, max_acc=1.0e6, std_epsilon=1e-8):
        self.max_acc = np.float32(max_acc)
        self.std_epsilon = np.float32(std_epsilon)
        self.acc_count = np.float32(0)
        self.num_acc = np.float32(0)
        self.acc_sum = np.zeros(dim, np.float32)
        self.acc_sum_sq = np.zeros(dim, np.float32)

    def update(self, x):
        x = np.asarray(x, np.float32)
        # sum(data; dims=2) in Julia's (F, M) == sum over rows of [M, F]; Float32 pairwise sum
        self.acc_sum = (self.acc_sum + x.sum(axis=0, dtype=np.float32)).astype(np.float32)
        self.acc_sum_sq = (self.acc_sum_sq + (x * x).sum(axis=0, dtype=np.float32)).astype(np.float32)
        self.acc_count = np.float32(self.acc_count + np.float32(x.shape[0]))
        self.num_acc = np.float32(self.num_acc + 1)

    def mean(self):
        return (self.acc_sum / max(self.acc_count, np.float32(1))).astype(np.float32)

    def std(self):
        c = max(self.acc_count, np.float32(1))
        m = self.mean()
        var = self.acc_sum_sq / c - m * m
        with np.errstate(invalid="ignore"):
            s = np.sqrt(var).astype(np.float32)
        s = np.where(np.isnan(s), self.std_epsilon, s)
        return np.maximum(s, self.std_epsilon).astype(np.float32)

    def __call__(self, x, accumulate=True):
        if accumulate and self.num_acc < self.max_acc:
            self.update(x)
        return ((np.asarray(x, np.float32) - self.mean()) / self.std()).astype(np.float32)

    def inverse(self, y):
        return (np.asarray(y, np.float32) * self.std() + self.mean()).astype(np.float32)


def inverse_data(norm, y):
    """GraphNetCore.inverse_data as called at src/solve.jl:207-209."""
    return norm.inverse(y)


# ----------------------------------------------------------------------------------------------
# Model: parameters
# ----------------------------------------------------------------------------------------------


@dataclass
class ModelConfig:
    """Mirrors GraphNetCore.build_model(quantities, dims, outputs, mps, layer_size,
    hidden_layers, device) reached through ``load`` at src/MeshGraphNets.jl:282-285."""
    node_in: int            # "quantities" (src/MeshGraphNets.jl:274)
    edge_in: int            # dims + 1 (src/graph.jl:51-52)
    out_dim: int            # "outputs" (src/MeshGraphNets.jl:277-280)
    latent: int = 128       # layer_size (src/MeshGraphNets.jl:37)
    mps: int = 15           # src/MeshGraphNets.jl:36
    hidden_layers: int = 2  # src/MeshGraphNets.jl:38
    ln_eps: float = 1e-5    # Lux 0.5 LayerNorm default epsilon (recalled)
    # the recalled internals as switches (mgn_model_config of include/mgn_b200.h)
    dense_layers: int = 0          # 0 = hidden_layers + 2
    ln_scale_first: bool = False   # flat order of the LayerNorm parameters: (bias, scale) recalled
    aggregate_post_residual: bool = False   # False: agg = scatter(+, m) (DeepMind order, recalled); True: scatter(+, ef + m)

    @property
    def n_dense(self):
        # recalled build_mlp: Dense(in,latent,relu), hidden_layers x Dense(latent,latent,relu),
        # Dense(latent,out)
        return self.dense_layers if self.dense_layers > 0 else self.hidden_layers + 2


@dataclass
class MlpSpec:
    name: str
    in_dim: int
    out_dim: int
    layer_norm: bool
    offset: int = 0
    # per dense layer: (w_off, b_off, in, out); LayerNorm: (bias_off, scale_off)
    dense: list = field(default_factory=list)
    ln: tuple = None
    size: int = 0


def mlp_specs(cfg: ModelConfig):
    """Flat Float32 parameter layout = the recalled ComponentArray order of
    Chain(encoder(node_model_fn, edge_model_fn), processors(edge_model_fn, node_model_fn)...,
    decoder(model)); each MLP = Dense(weight, bias)... then LayerNorm(bias, scale).
    Dense weight is Julia (out x in) column-major == C row-major [in][out]."""
    D, L = cfg.latent, cfg.n_dense
    specs = [MlpSpec("encoder.node", cfg.node_in, D, True), MlpSpec("encoder.edge", cfg.edge_in, D, True)]
    for k in range(cfg.mps):
        specs.append(MlpSpec(f"processor{k + 1}.edge", 3 * D, D, True))
        specs.append(MlpSpec(f"processor{k + 1}.node", 2 * D, D, True))
    specs.append(MlpSpec("decoder", D, cfg.out_dim, False))
    off = 0
    for s in specs:
        s.offset = off
        dims = [s.in_dim] + [D] * (L - 1) + [s.out_dim]
        for l in range(L):
            i, o = dims[l], dims[l + 1]
            s.dense.append((off, off + i * o, i, o))
            off += i * o + o
        if s.layer_norm:   # s.ln = (bias offset, scale offset)
            s.ln = (off + s.out_dim, off) if cfg.ln_scale_first else (off, off + s.out_dim)
            off += 2 * s.out_dim
        s.size = off - s.offset
    return specs, off


def init_params(cfg: ModelConfig, seed=1234, dtype=np.float32):
    """Glorot-uniform weights U(+-sqrt(6/(in+out))), zero bias, LayerNorm scale 1 / bias 0
    (Lux defaults).  Julia's RNG stream is NOT reproduced - tensors are exchanged."""
    specs, P = mlp_specs(cfg)
    rng = np.random.Generator(np.random.PCG64(seed))
    p = np.zeros(P, dtype=np.float64)
    for s in specs:
        for (w, b, i, o) in s.dense:
            lim = np.sqrt(6.0 / (i + o))
            p[w:w + i * o] = rng.uniform(-lim, lim, size=i * o)
        if s.ln is not None:
            p[s.ln[1]:s.ln[1] + s.out_dim] = 1.0
    return p.astype(dtype)


# ----------------------------------------------------------------------------------------------
# Model: forward / backward (explicit, any float dtype)
# ----------------------------------------------------------------------------------------------


def _mlp_forward(p, s: MlpSpec, x, eps, tape=None):
    """Lux.Dense chain + Lux.LayerNorm: y = W x + b, relu on all but the last
    Dense; LayerNorm over the feature dim with biased variance, y = xhat*scale + bias."""
    h = x
    hs = [x]
    L = len(s.dense)
    for l, (w, b, i, o) in enumerate(s.dense):
        W = p[w:w + i * o].reshape(i, o)
        h = h @ W + p[b:b + o]
        if l < L - 1:
            h = np.maximum(h, 0)
        hs.append(h)
    y = h
    if s.ln is not None:
        mu = y.mean(axis=1, keepdims=True)
        var = ((y - mu) ** 2).mean(axis=1, keepdims=True)
        rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
        xhat = (y - mu) * rstd
        out = xhat * p[s.ln[1]:s.ln[1] + s.out_dim] + p[s.ln[0]:s.ln[0] + s.out_dim]
    else:
        xhat = rstd = None
        out = y
    if tape is not None:
        tape.append((s, hs, xhat, rstd))
    return out


def _mlp_backward(p, g, rec, dout):
    """Reverse of _mlp_forward; accumulates into flat gradient ``g``; returns d(input)."""
    s, hs, xhat, rstd = rec
    L = len(s.dense)
    if s.ln is not None:
        scale = p[s.ln[1]:s.ln[1] + s.out_dim]
        g[s.ln[0]:s.ln[0] + s.out_dim] += dout.sum(axis=0)
        g[s.ln[1]:s.ln[1] + s.out_dim] += (dout * xhat).sum(axis=0)
        dxh = dout * scale
        dz = rstd * (dxh - dxh.mean(axis=1, keepdims=True) - xhat * (dxh * xhat).mean(axis=1, keepdims=True))
    else:
        dz = dout
    for l in range(L - 1, -1, -1):
        w, b, i, o = s.dense[l]
        W = p[w:w + i * o].reshape(i, o)
        if l < L - 1:
            dz = dz * (hs[l + 1] > 0)
        g[w:w + i * o] += (hs[l].T @ dz).reshape(-1)
        g[b:b + o] += dz.sum(axis=0)
        dz = dz @ W.T
    return dz


def scatter_add(m, receivers0, n_nodes):
    """NNlib.scatter(+, m, receivers) on CPU: sequential adds in ascending edge id."""
    agg = np.zeros((n_nodes, m.shape[1]), dtype=m.dtype)
    np.add.at(agg, receivers0, m)
    return agg


def model_forward(cfg: ModelConfig, params, nf, ef, senders, receivers, index_base=1, tape=None,
                  dtype=None):
    """``mgn.model(graph, ps, st)`` as called at src/solve.jl:200 / inside step!
    (src/strategies.jl:421): Encoder -> mps x Processor -> Decoder.
    Processor (recalled, DeepMind order): m = LN(MLP_e([nf[s]; nf[r]; ef])),
    agg = scatter(+, m, r) (pre-residual), n = LN(MLP_n([nf; agg])), nf += n, ef += m.
    ``cfg.aggregate_post_residual`` is the other reading of the block: agg = scatter(+, ef + m, r)."""
    dtype = dtype or params.dtype
    p = params.astype(dtype, copy=False)
    nf = np.asarray(nf, dtype=dtype)
    ef = np.asarray(ef, dtype=dtype)
    s0 = np.asarray(senders, np.int64) - index_base
    r0 = np.asarray(receivers, np.int64) - index_base
    specs, _ = mlp_specs(cfg)
    N = nf.shape[0]
    x = _mlp_forward(p, specs[0], nf, cfg.ln_eps, tape)
    e = _mlp_forward(p, specs[1], ef, cfg.ln_eps, tape)
    for k in range(cfg.mps):
        se, sn = specs[2 + 2 * k], specs[3 + 2 * k]
        m = _mlp_forward(p, se, np.concatenate([x[s0], x[r0], e], axis=1), cfg.ln_eps, tape)
        agg = scatter_add(e + m if cfg.aggregate_post_residual else m, r0, N)
        n = _mlp_forward(p, sn, np.concatenate([x, agg], axis=1), cfg.ln_eps, tape)
        x = x + n
        e = e + m
    return _mlp_forward(p, specs[-1], x, cfg.ln_eps, tape)


def model_backward(cfg: ModelConfig, params, tape, dout, senders, receivers, n_nodes, index_base=1):
    """Reverse-mode of model_forward (what Zygote derives for step! / the ZygoteVJP of
    ode_step, src/strategies.jl:183-194): returns (d_params[P], d_nf_in[N, node_in])."""
    p = params.astype(dout.dtype, copy=False)
    s0 = np.asarray(senders, np.int64) - index_base
    r0 = np.asarray(receivers, np.int64) - index_base
    D = cfg.latent
    g = np.zeros_like(p)
    recs = list(tape)
    dx = _mlp_backward(p, g, recs.pop(), dout)                 # decoder
    de = np.zeros((s0.shape[0], D), dtype=dout.dtype)
    for k in range(cfg.mps - 1, -1, -1):
        rec_n = recs.pop()
        rec_e = recs.pop()
        din = _mlp_backward(p, g, rec_n, dx)                   # nf' = nf + n
        dx = dx + din[:, :D]
        dagg = din[:, D:]
        dm = de + dagg[r0]                                     # ef' = ef + m ; agg = scatter(m)
        if cfg.aggregate_post_residual:                        # agg = scatter(ef'): the aggregation adjoint also flows
            de = dm                                            # down the residual path
        din = _mlp_backward(p, g, rec_e, dm)
        np.add.at(dx, s0, din[:, :D])
        np.add.at(dx, r0, din[:, D:2 * D])
        de = de + din[:, 2 * D:]
    rec_e = recs.pop()
    rec_n = recs.pop()
    _mlp_backward(p, g, rec_e, de)
    dnf = _mlp_backward(p, g, rec_n, dx)
    return g, dnf


def mse_reduce(target, out):
    """GraphNetCore.mse_reduce (passed at src/strategies.jl:421): sum over feature rows of
    (target - out)^2 -> one value per node."""
    return ((target - out) ** 2).sum(axis=1)


def loss_and_dout(out, target, mask, index_base=1):
    """step! loss (recalled): mean(mse_reduce(target, out)[mask]); returns (loss, dloss/dout)."""
    m0 = np.asarray(mask, np.int64) - index_base
    err = mse_reduce(target, out)
    loss = err[m0].mean()
    dout = np.zeros_like(out)
    dout[m0] = 2.0 * (out[m0] - target[m0]) / out.dtype.type(m0.shape[0])
    return loss, dout


def step(cfg, params, nf, ef, senders, receivers, target, mask, index_base=1, dtype=None):
    """GraphNetCore.step!(mgn, graph, target, mask, mse_reduce) -> (gs, loss) as called at
    src/strategies.jl:421."""
    dtype = dtype or params.dtype
    tape = []
    out = model_forward(cfg, params, nf, ef, senders, receivers, index_base, tape, dtype)
    loss, dout = loss_and_dout(out, np.asarray(target, dtype), mask, index_base)
    g, dnf = model_backward(cfg, params.astype(dtype), tape, dout, senders, receivers, nf.shape[0], index_base)
    return g, loss, out, dnf


def adam_update(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """Optimisers.Adam (0.3) update used at src/MeshGraphNets.jl:374-378 (recalled rule):
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)."""
    f = np.float32
    m = (f(b1) * m + f(1 - b1) * g).astype(np.float32)
    v = (f(b2) * v + f(1 - b2) * g * g).astype(np.float32)
    mh = m / f(1 - b1 ** t)
    vh = v / f(1 - b2 ** t)
    p = (p - mh / (np.sqrt(vh) + f(eps)) * f(lr)).astype(np.float32)
    return p, m, v


# ----------------------------------------------------------------------------------------------
# Callers restated: build_graph / ode_step / rollout
# ----------------------------------------------------------------------------------------------


def build_graph_features(n_norms, e_norm, field_values, fields, node_type_onehot, edge_feats):
    """src/graph.jl:75-97: nf = vcat(n_norm[f](data[f]) for f in fields..., n_norm["node_type"](onehot))
    (fields first, node type LAST); ef = e_norm(edge_features).  Note the node_type normaliser
    is evaluated first (graph.jl:80) - this matters for the order online statistics accumulate."""
    nt = n_norms["node_type"](node_type_onehot)
    cols = [n_norms[f](field_values[f]) for f in fields]
    nf = np.concatenate(cols + [nt], axis=1).astype(np.float32)
    ef = e_norm(edge_feats)
    return nf, ef


def ode_step(cfg, params, x, n_norms, e_norm, o_norms, fields, target_fields, target_dims,
             inputs, node_type_onehot, edge_feats, senders, receivers, vmask, dtype=np.float32):
    """src/solve.jl:188-219: split the state by target field, build_graph(datapoint=1), model
    forward, inverse_data per target field (solve.jl:205-210), times val_mask (solve.jl:218).
    ``x`` is [N, sum(target_dims)]."""
    vals = dict(inputs)
    off = 0
    for k, d in zip(target_fields, target_dims):
        vals[k] = x[:, off:off + d]
        off += d
    nf, ef = build_graph_features(n_norms, e_norm, vals, fields, node_type_onehot, edge_feats)
    out = model_forward(cfg, params, nf, ef, senders, receivers, 1, None, dtype)
    buf = np.empty_like(out)
    off = 0
    for k, d in zip(target_fields, target_dims):
        buf[:, off:off + d] = inverse_data(o_norms[k], out[:, off:off + d].astype(np.float32))
        off += d
    return (buf * vmask).astype(x.dtype)


def inflow_index(t, saves_dt):
    """src/solve.jl:151: floor(Int, t / saves_dt) + 1 in Float32 (returned 0-based)."""
    return int(np.floor(np.float32(t) / np.float32(saves_dt)))


def rollout_euler(rhs, x0, saves, dt, inflow_fn=None):
    """src/solve.jl:42-68 with ``solve(prob, Euler(); adaptive=false, dt=dt, saveat=saves)``
    (the examples/cylinder_flow/cylinder_flow.jl:79-84 configuration): returns the saved
    states at ``saves``.  ``inflow_fn(x, t)`` applies the in-place overwrite of src/solve.jl:151."""
    x = np.array(x0, copy=True)
    sol = [x.copy()]
    for i in range(len(saves) - 1):
        t = np.float32(saves[i])
        if inflow_fn is not None:
            x = inflow_fn(x, t)
        x = (x + np.float32(dt) * rhs(x, t)).astype(x0.dtype)
        sol.append(x.copy())
    return sol


# ----------------------------------------------------------------------------------------------
# Synthetic workloads
# ----------------------------------------------------------------------------------------------


def cylinder_flow_mesh(nx=65, ny=29, lx=1.6, ly=0.41):
    """CylinderFlow-shaped structured triangulated grid: N = nx*ny = 1885,
    C = 2*(nx-1)*(ny-1) = 3584 triangles (0-based Int32, same diagonal), node types
    x=0 -> 4, x=max -> 5, y=0 / y=max -> 6, interior 0, plus one interior block of type 1 so
    that inflow_mask (src/MeshGraphNets.jl:428) is exercised."""
    xs = np.linspace(0.0, lx, nx, dtype=np.float32)
    ys = np.linspace(0.0, ly, ny, dtype=np.float32)
    pos = np.stack(np.meshgrid(xs, ys, indexing="ij"), axis=-1).reshape(-1, 2).astype(np.float32)
    idx = np.arange(nx * ny, dtype=np.int32).reshape(nx, ny)
    a = idx[:-1, :-1].reshape(-1)
    b = idx[1:, :-1].reshape(-1)
    c = idx[1:, 1:].reshape(-1)
    d = idx[:-1, 1:].reshape(-1)
    cells = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=0).astype(np.int32)
    nt = np.zeros((nx, ny), dtype=np.int32)
    nt[:, 0] = 6
    nt[:, -1] = 6
    nt[0, :] = 4
    nt[-1, :] = 5
    nt[1, 1:-1] = 1
    return pos, cells, nt.reshape(-1)


def chain_mesh(n):
    """1-D chain (config 4): positions on [0,1], edges from create_edges_1d."""
    pos = np.linspace(0.0, 1.0, n, dtype=np.float32)[:, None]
    nt = np.zeros(n, dtype=np.int32)
    nt[0] = 4
    nt[-1] = 5
    return pos, create_edges_1d(n), nt


def synthetic_velocity(pos, T, seed=1234, noise=0.1):
    """Smooth field + N(0, noise^2), [T, N, 2] Float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(T, dtype=np.float32)[:, None] * 0.01
    x, y = pos[None, :, 0], pos[None, :, -1]
    u = np.sin(2 * np.pi * (x / 1.6 - t)) * np.cos(np.pi * y / 0.41) + 1.0
    v = 0.3 * np.cos(2 * np.pi * (x / 1.6 + t)) * np.sin(np.pi * y / 0.41)
    vel = np.stack([u, v], axis=-1)
    return (vel + rng.normal(0.0, noise, size=vel.shape)).astype(np.float32)
