"""torch-CPU restatement of the reference's op sequence, used (a) in fp64 with autograd to
cross-check the hand-written backward of oracle/mgn_oracle.py and (b) in fp32 on all host cores
(MKL sgemm) as the timed CPU baseline of bench.py.

TEST INFRASTRUCTURE ONLY - never imported by the product package.  PARITY UNPINNED (see
oracle/mgn_oracle.py).  The op sequence follows what the Julia reference executes on CPU:
gather, gather, vcat, L x (Dense: sgemm + bias + relu), LayerNorm, sequential scatter-add,
vcat, node MLP, LayerNorm, residual adds, reverse-mode autograd (Zygote's role),
Adam (Optimisers.update, src/MeshGraphNets.jl:374-378)."""
from __future__ import annotations

import torch

from mgn_oracle import ModelConfig, mlp_specs


def _mlp(p, s, x, eps):
    L = len(s.dense)
    h = x
    for l, (w, b, i, o) in enumerate(s.dense):
        h = h @ p[w:w + i * o].view(i, o) + p[b:b + o]
        if l < L - 1:
            h = torch.relu(h)
    if s.ln is not None:
        mu = h.mean(dim=1, keepdim=True)
        var = ((h - mu) ** 2).mean(dim=1, keepdim=True)
        h = (h - mu) / torch.sqrt(var + eps) * p[s.ln[1]:s.ln[1] + s.out_dim] + p[s.ln[0]:s.ln[0] + s.out_dim]
    return h


def model_forward(cfg: ModelConfig, p, nf, ef, s0, r0):
    """Same structure as mgn_oracle.model_forward; s0/r0 are 0-based int64 tensors."""
    specs, _ = mlp_specs(cfg)
    x = _mlp(p, specs[0], nf, cfg.ln_eps)
    e = _mlp(p, specs[1], ef, cfg.ln_eps)
    for k in range(cfg.mps):
        m = _mlp(p, specs[2 + 2 * k], torch.cat([x[s0], x[r0], e], dim=1), cfg.ln_eps)
        agg = torch.zeros_like(x).index_add_(0, r0, m)
        n = _mlp(p, specs[3 + 2 * k], torch.cat([x, agg], dim=1), cfg.ln_eps)
        x = x + n
        e = e + m
    return _mlp(p, specs[-1], x, cfg.ln_eps)


def loss_fn(out, target, m0):
    return ((target - out) ** 2).sum(dim=1)[m0].mean()


def train_step(cfg, p, opt_state, nf, ef, s0, r0, target, m0, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """One derivative-training step: forward, masked MSE, backward, Adam.  Returns the loss."""
    p.requires_grad_(True)
    if p.grad is not None:
        p.grad = None
    loss = loss_fn(model_forward(cfg, p, nf, ef, s0, r0), target, m0)
    loss.backward()
    with torch.no_grad():
        g = p.grad
        opt_state["t"] += 1
        t = opt_state["t"]
        opt_state["m"].mul_(b1).add_(g, alpha=1 - b1)
        opt_state["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        mh = opt_state["m"] / (1 - b1 ** t)
        vh = opt_state["v"] / (1 - b2 ** t)
        p.sub_(mh / (vh.sqrt() + eps) * lr)
    return float(loss)
