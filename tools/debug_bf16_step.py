"""Per-tensor gradient error of the bf16 (tcgen05) mode against the fp64 oracle (debug aid)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import mgn_oracle as orc  # noqa: E402
import mgn_pkg  # noqa: E402
from test_gpu_tc_parity import _problem, dev, rel  # noqa: E402

pkg = mgn_pkg.pkg
nx, ny, mps, hidden = [int(a) for a in (sys.argv[1:5] or (12, 9, 3, 2))]
cfg, ps, nf, ef, s, r, tgt, mask = _problem(nx, ny, mps, hidden=hidden)
g_o, loss_o, out_o, dnf_o = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
model = pkg.Model(9, 3, 2, mps, 128, hidden, compute_mode=pkg.COMPUTE_BF16)
graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
mgn = pkg.GraphNetwork(model, dev(ps), None, None, None, None)
(gs,), loss = pkg.step_(mgn, graph, dev(tgt), dev(mask))
torch.cuda.synchronize()
print("loss", float(loss.cpu()), "oracle", loss_o, "grad rel", rel(gs.cpu().numpy(), g_o))
for name, off, rows, cols in model.param_layout():
    ref = g_o[off:off + rows * cols]
    got = gs[off:off + rows * cols].cpu().numpy()
    e = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    flag = " <<<" if e > 0.1 else ""
    print(f"{name:44s} |ref| {np.linalg.norm(ref):10.3e}  |got| {np.linalg.norm(got):10.3e}  rel {e:9.2e}{flag}")
out = model.forward(graph, dev(ps), training=True)
_, dout_o = orc.loss_and_dout(out_o, tgt.astype(np.float64), mask)
dps, dnf = model.backward(graph, dev(ps), dev(dout_o.astype(np.float32)), want_dnf=True)
print("dnf rel", rel(dnf.cpu().numpy(), dnf_o))
