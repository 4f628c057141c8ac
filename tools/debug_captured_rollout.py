#!/usr/bin/env python
"""Diagnostic: eager rollout vs eager rollout vs CapturedRollout replays, per saved step."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import mgn_pkg  # noqa: E402
import mgn_oracle as orc  # noqa: E402
from test_gpu_callers import _setup, dev  # noqa: E402

pkg = mgn_pkg.pkg
for mode in (0, 1):
    data_h, data, meta, mgn, (node_type, senders, receivers, ef), o = _setup(pkg, T=12, mode=mode)
    x0 = data_h["velocity"][0]
    for n_g, x in ((mgn.n_norm["velocity"], x0), (mgn.e_norm, o["ef"]),
                   (mgn.o_norm["velocity"], (data_h["velocity"][1] - x0) / np.float32(0.01))):
        n_g(dev(x))
        n_g.max_acc = 0.0
    vm = dev(orc.val_mask(o["nt"], [0, 5], 2))
    inflow = dev(np.repeat((o["nt"] == 1)[:, None], 2, axis=1))
    saves = [np.float32(0.01) * i for i in range(6)]
    args = (mgn, {"velocity": dev(x0)}, ["velocity"], meta, ["velocity"], {"velocity": 2}, node_type, ef, senders,
            receivers, vm, inflow, data, 0.0, 0.05, 0.01, saves)
    st0 = [n.state.clone() for n in (mgn.n_norm["velocity"], mgn.e_norm, mgn.o_norm["velocity"])]
    e1, _ = pkg.rollout(*args, solver="euler")
    e2, _ = pkg.rollout(*args, solver="euler")
    cap = pkg.CapturedRollout(*args, solver="euler")
    g1 = [t.clone() for t in cap.replay()[0]]
    g2 = [t.clone() for t in cap.replay()[0]]
    e3, _ = pkg.rollout(*args, solver="euler")
    st1 = [n.state.clone() for n in (mgn.n_norm["velocity"], mgn.e_norm, mgn.o_norm["velocity"])]
    print("mode", mode, "normaliser state unchanged:", [bool(torch.equal(a, b)) for a, b in zip(st0, st1)])
    for name, a, b in (("eager1-eager2", e1, e2), ("eager1-graph1", e1, g1), ("graph1-graph2", g1, g2), ("eager1-eager3", e1, e3)):
        print(" ", name, ["%.2e" % float((x - y).abs().max()) for x, y in zip(a, b)])
    print("  |x| max", float(e1[-1].abs().max()))
