"""Localise kernel-vs-bf16-model forward differences (debug aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import mgn_oracle as orc  # noqa: E402
import mgn_oracle_bf16 as ob  # noqa: E402
import mgn_pkg  # noqa: E402
from test_gpu_tc_parity import _problem, dev, rel  # noqa: E402

pkg = mgn_pkg.pkg
for (nx, ny, mps, hidden) in [(12, 9, 0, 2), (12, 9, 1, 2), (12, 9, 3, 2), (12, 9, 1, 0)]:
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(nx, ny, mps, hidden=hidden)
    out_b = ob.forward_bf16(cfg, ps, nf, ef, s, r)
    model = pkg.Model(9, 3, 2, mps, 128, hidden, compute_mode=pkg.COMPUTE_BF16)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    out = model.forward(graph, dev(ps), training=False).cpu().numpy()
    err = np.abs(out - out_b).max(axis=1) / np.abs(out_b).max()
    deg = np.bincount(r - 1, minlength=nf.shape[0])
    print(f"mps={mps} hidden={hidden} rel={rel(out, out_b):.2e}  max node err {err.max():.2e}  "
          f"nodes with err>1e-4: {(err > 1e-4).sum()}/{len(err)}")
    for d in sorted(set(deg)):
        m = deg == d
        print(f"   in-degree {d}: {m.sum():3d} nodes, mean err {err[m].mean():.2e}, exact-ish {(err[m] < 1e-6).sum()}")
