"""CPU model of the product's MGN_COMPUTE_BF16 arithmetic (tcgen05 path): the SAME algorithm as
oracle/mgn_oracle.py (which restates the reference), with a round-to-nearest-even to bfloat16
inserted at exactly the points where the CUDA kernels store a bf16 value (GEMM operands, saved
activations, staged gradient tiles, and the edge latent / its gradient, which have no fp32 master copy).  Accumulation is exact (float64) where the kernels accumulate
in fp32.

TEST INFRASTRUCTURE ONLY (same rule as mgn_oracle.py).  Purpose: split the bf16-mode parity claim
into two checkable halves -
  (1) kernels == this model to ~1e-3 (only fp32-vs-exact accumulation order differs), and
  (2) this model vs the fp64 restatement of the reference = the cost of bf16 operands, a property
      of the number format that is stated in DESIGN.md and measured in tests/test_oracle_bf16.py.
Rounding points follow csrc/tc_kernels.cu (forward) and csrc/tc_bwd_kernels.cu (backward).
"""
from __future__ import annotations

import numpy as np

import mgn_oracle as orc


def q(x):
    """Round to nearest-even bfloat16, returned as float64 holding the bf16 value."""
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape).astype(np.float64)


def f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


class _Mlp:
    """Forward of one MLP with the kernel's rounding; keeps what the backward needs."""

    def __init__(self, p, spec, eps):
        self.s = spec
        self.p = p
        self.eps = np.float32(eps)
        self.L = len(spec.dense)

    def W(self, l):
        w, b, i, o = self.s.dense[l]
        return self.p[w:w + i * o].reshape(i, o)

    def b(self, l):
        w, b, i, o = self.s.dense[l]
        return self.p[b:b + o]

    def forward(self, xq):
        """xq: bf16-valued input.  Returns the fp32 LayerNorm output m (or the linear output)."""
        self.x = xq
        self.h = []
        a = xq
        for l in range(self.L):
            z = f32(a @ q(self.W(l))) + self.b(l)
            z = f32(z)
            if l < self.L - 1:
                a = q(np.maximum(z, 0))
                self.h.append(a)
        if self.s.ln is None:
            return z
        # LayerNorm statistics exactly as the kernel computes them: one pass, data shifted by the row's first
        # element, sequential fp32 sums (fused multiply-add for the squares) over the left and the right half of the
        # row separately (two threads share a row), halves added, biased variance
        f = np.float32
        z32 = z.astype(f)
        shift = z32[:, :1]
        d = (z32 - shift).astype(f)
        half = z32.shape[1] // 2
        parts = []
        for lo in (0, half):
            ssum = np.zeros(z32.shape[0], f)
            qsum = np.zeros(z32.shape[0], f)
            for j in range(lo, lo + half):
                ssum = (ssum + d[:, j]).astype(f)
                qsum = (d[:, j].astype(np.float64) * d[:, j] + qsum).astype(f)
            parts.append((ssum, qsum))
        ssum = (parts[0][0] + parts[1][0]).astype(f)
        qsum = (parts[0][1] + parts[1][1]).astype(f)
        ms = (ssum * f(1.0 / 128.0)).astype(f)
        mean32 = (shift[:, 0] + ms).astype(f)
        var128 = (np.maximum((qsum * f(1.0 / 128.0)).astype(f) - (ms * ms).astype(f), f(0)) * f(128)).astype(f)
        rstd32 = (f(1) / np.sqrt(((var128 * f(1.0 / 128.0)).astype(f) + self.eps).astype(f))).astype(f)
        self.rstd = rstd32.astype(np.float64)[:, None]
        self.xhat = q(((z32 - mean32[:, None]).astype(f) * rstd32[:, None]).astype(f))
        sc = self.p[self.s.ln[1]:self.s.ln[1] + 128]
        bi = self.p[self.s.ln[0]:self.s.ln[0] + 128]
        return f32(self.xhat * sc + bi)

    # ---- backward pieces
    def head_ln(self, g, dy):
        sc = self.p[self.s.ln[1]:self.s.ln[1] + 128]
        g[self.s.ln[0]:self.s.ln[0] + 128] += dy.sum(axis=0)
        g[self.s.ln[1]:self.s.ln[1] + 128] += (dy * self.xhat).sum(axis=0)
        dxh = dy * sc
        m1 = dxh.sum(axis=1, keepdims=True) / 128.0
        m2 = (dxh * self.xhat).sum(axis=1, keepdims=True) / 128.0
        return q(self.rstd * (dxh - m1 - self.xhat * m2))

    def chain(self, g, z, top):
        """Layers top .. 1: dW, db, dZ of the layer below.  Returns dZ_0."""
        w, b, i, o = self.s.dense[top]
        g[b:b + o] += z.sum(axis=0)
        for l in range(top, 0, -1):
            w, b, i, o = self.s.dense[l]
            g[w:w + i * o] += (self.h[l - 1].T @ z).reshape(-1)
            dx = f32(z @ q(self.W(l)).T)
            z = q(np.where(self.h[l - 1] > 0, dx, 0.0))
            wb, bb, ib, ob = self.s.dense[l - 1]
            g[bb:bb + ob] += z.sum(axis=0)
        return z

    def input_dx(self, g, z0):
        """First layer with a bf16 block input: dW_0 and the staged (bf16) dX."""
        w, b, i, o = self.s.dense[0]
        g[w:w + i * o] += (self.x.T @ z0).reshape(-1)
        return q(f32(z0 @ q(self.W(0)).T))


def _seg_sum(rows, order, keys, n_nodes, init=None):
    """Sequential sum of rows[order[j]] into node keys[order[j]] (fp32 running sum)."""
    out = np.zeros((n_nodes, rows.shape[1])) if init is None else np.array(init, dtype=np.float64)
    out = out.astype(np.float32)
    for j in order:
        out[keys[j]] = out[keys[j]] + rows[j].astype(np.float32)
    return out.astype(np.float64)


def step_bf16(cfg: orc.ModelConfig, params, nf, ef, senders, receivers, target, mask, index_base=1):
    """GraphNetCore.step! in the tensor-core mode's arithmetic -> (grads, loss, out, d_nf_raw)."""
    assert cfg.latent == 128
    p = f32(params)
    s0 = np.asarray(senders, np.int64) - index_base
    r0 = np.asarray(receivers, np.int64) - index_base
    N, E = nf.shape[0], s0.shape[0]
    specs, P = orc.mlp_specs(cfg)
    perm = np.argsort(r0, kind="stable")           # CSR order: stable by receiver
    sc_, rc_ = s0[perm], r0[perm]                  # sender / receiver of each CSR slot
    inv_perm = np.empty(E, np.int64)
    inv_perm[perm] = np.arange(E)
    csc = inv_perm[np.argsort(s0, kind="stable")]  # CSR slots in CSC order (stable by sender over original ids)
    mk = lambda k: _Mlp(p, specs[k], cfg.ln_eps)

    enc_n, enc_e = mk(0), mk(1)
    nf32 = enc_n.forward(q(nf))
    ef32 = q(enc_e.forward(q(np.asarray(ef)[perm])))   # the edge latent is STORED in bf16 only (no fp32 master)
    nf16, ef16 = [q(nf32)], [ef32]
    agg16, edges, nodes = [], [], []
    for k in range(cfg.mps):
        me, mn = mk(2 + 2 * k), mk(3 + 2 * k)
        m = me.forward(np.concatenate([nf16[k][sc_], nf16[k][rc_], ef16[k]], axis=1))
        ef32 = q(ef32 + m)                             # residual on the bf16 latent, rounded once
        agg = q(_seg_sum(m, range(E), rc_, N))
        n = mn.forward(np.concatenate([nf16[k], agg], axis=1))
        nf32 = f32(nf32 + n)
        nf16.append(q(nf32))
        ef16.append(q(ef32))
        agg16.append(agg)
        edges.append(me)
        nodes.append(mn)
    dec = mk(len(specs) - 1)
    out = dec.forward(nf16[-1])
    loss, dout = orc.loss_and_dout(out, np.asarray(target, np.float64), mask, index_base)
    dout = f32(dout)

    # ---------------- backward
    g = np.zeros(P)
    L = cfg.n_dense
    w, b, i, o = dec.s.dense[L - 1]
    hq = dec.h[L - 2]
    g[w:w + i * o] += (hq.T @ dout).reshape(-1)
    g[b:b + o] += dout.sum(axis=0)
    z = q(np.where(hq > 0, f32(dout @ p[w:w + i * o].reshape(i, o).T), 0.0))
    wb, bb, ib, ob = dec.s.dense[L - 2]
    g[bb:bb + ob] += z.sum(axis=0)
    for l in range(L - 2, 0, -1):                  # chain without the head's db (added above)
        wl, bl, il, ol = dec.s.dense[l]
        g[wl:wl + il * ol] += (dec.h[l - 1].T @ z).reshape(-1)
        z = q(np.where(dec.h[l - 1] > 0, f32(z @ q(dec.W(l)).T), 0.0))
        wq, bq, iq, oq = dec.s.dense[l - 1]
        g[bq:bq + oq] += z.sum(axis=0)
    d_nf = f32(dec.input_dx(g, z))
    d_ef = None
    for k in range(cfg.mps - 1, -1, -1):
        mn, me = nodes[k], edges[k]
        z0 = mn.chain(g, mn.head_ln(g, d_nf), L - 1)
        dx = mn.input_dx(g, z0)
        d_nf = f32(d_nf + dx[:, :128])
        d_agg = f32(dx[:, 128:])
        dy = d_agg[rc_] if d_ef is None else f32(d_ef + d_agg[rc_])
        z0 = me.chain(g, me.head_ln(g, dy), L - 1)
        dx = me.input_dx(g, z0)
        recv = _seg_sum(dx[:, 128:256], range(E), rc_, N)       # tile-local segmented sum, stored ...
        d_ef = dx[:, 256:] if d_ef is None else q(d_ef + dx[:, 256:])   # gradient of the edge latent: bf16 images too
        d_nf = _seg_sum(dx[:, :128], csc, sc_, N, init=f32(d_nf + recv))   # ... then added with the sender rows
    d_raw = None
    for enc, dy, raw in ((enc_e, d_ef, np.asarray(ef)[perm]), (enc_n, d_nf, nf)):
        if dy is None:
            continue
        z0 = enc.chain(g, enc.head_ln(g, dy), L - 1)
        w, b, i, o = enc.s.dense[0]
        g[w:w + i * o] += (f32(raw).T @ z0).reshape(-1)
        if enc is enc_n:
            d_raw = z0 @ p[w:w + i * o].reshape(i, o).T
    return g, loss, out, d_raw


def forward_bf16(cfg, params, nf, ef, senders, receivers, index_base=1):
    """mgn.model(graph, ps, st) in the tensor-core mode's arithmetic."""
    n = nf.shape[0]
    g, loss, out, _ = step_bf16(cfg, params, nf, ef, senders, receivers, np.zeros((n, cfg.out_dim)),
                                np.arange(index_base, index_base + n), index_base)
    return out
