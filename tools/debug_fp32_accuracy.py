#!/usr/bin/env python
"""Diagnostic: where does the fp32 (CUDA-core) mode lose digits against the fp64 oracle?  Per-tensor relative errors of
the parameter gradient for the GPU fp32 mode and for the numpy fp32 run of the oracle (same arithmetic width)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import mgn_pkg  # noqa: E402
import mgn_oracle as orc  # noqa: E402

pkg = mgn_pkg.pkg
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300))
rng = np.random.default_rng(0)
pos, cells, nt = orc.cylinder_flow_mesh(12, 9)
s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
N, E = pos.shape[0], s.shape[0]
for mps in (0, 1, 3):
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, 2)
    ps = orc.init_params(cfg, seed=3)
    nf = rng.normal(size=(N, 9)).astype(np.float32)
    ef = rng.normal(size=(E, 3)).astype(np.float32)
    tgt = rng.normal(size=(N, 2)).astype(np.float32)
    mask = orc.node_mask(nt, [0, 5])
    g64, l64, o64, d64 = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    g32, l32, o32, d32 = orc.step(cfg, ps, nf, ef, s, r, tgt, mask, dtype=np.float32)
    model = pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=pkg.COMPUTE_FP32)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    out = model.forward(graph, dev(ps), training=True).cpu().numpy()
    _, dout = orc.loss_and_dout(o64, tgt.astype(np.float64), mask)
    dps, dnf = model.backward(graph, dev(ps), dev(dout.astype(np.float32)), want_dnf=True)
    dps, dnf = dps.cpu().numpy(), dnf.cpu().numpy()
    print(f"mps={mps}: out gpu {rel(out, o64):.2e} np32 {rel(o32, o64):.2e} | grad gpu {rel(dps, g64):.2e} np32 "
          f"{rel(g32, g64):.2e} | dnf gpu {rel(dnf, d64):.2e} np32 {rel(d32, d64):.2e}")
    rows = []
    for name, off, rows_, cols in model.param_layout():
        n = rows_ * cols
        a, b, c = dps[off:off + n], g64[off:off + n], g32[off:off + n]
        rows.append((rel(a, b), rel(c, b), name, float(np.linalg.norm(b))))
    for e_gpu, e_np, name, nrm in sorted(rows, reverse=True)[:8]:
        print(f"    {name:40s} gpu {e_gpu:.2e}  np32 {e_np:.2e}  |g| {nrm:.2e}")
