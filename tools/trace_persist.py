"""Stage timeline of the persistent forward kernel (debug build: python meshgraphnets.jl_b200/build.py --trace):
per stage, CTA 0's time in the stage body and in the grid barrier behind it.  Usage: python tools/trace_persist.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
os.environ["MGN_FWD_PERSIST"] = "2"
import mgn_oracle as orc  # noqa: E402
import mgn_pkg  # noqa: E402

pkg = mgn_pkg.pkg
pos, cells, nt = orc.cylinder_flow_mesh(65, 29)
s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
N, E = pos.shape[0], s.shape[0]
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
nf = dev(rng.normal(size=(N, 9)).astype(np.float32))
ef = dev(rng.normal(size=(E, 3)).astype(np.float32))
model, ps, _ = pkg.build_model(9, 2, 2, 15, 128, 2, compute_mode=pkg.COMPUTE_BF16)
graph = pkg.FeatureGraph(nf, ef, dev(s), dev(r))
lib = pkg.load()
lib.mgn_debug_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
for _ in range(3):
    model.forward(graph, ps, training=False)
torch.cuda.synchronize()
buf = torch.zeros(4 * 512, dtype=torch.int64, device="cuda")
lib.mgn_debug_trace(buf.data_ptr(), 0, 0)
model.forward(graph, ps, training=False)
torch.cuda.synchronize()
t = buf.cpu().numpy()[:3 * 33].reshape(33, 3).astype(np.float64)
t0 = t[0, 0]
print("stage  top_us  body_us  barrier_us")
for i in range(33):
    nxt = t[i, 2] if t[i, 2] > 0 else t[i, 1]
    print(f"{i:3d} {(t[i, 0] - t0) / 1e3:8.2f} {(t[i, 1] - t[i, 0]) / 1e3:7.2f} {(nxt - t[i, 1]) / 1e3:7.2f}")
print("total us", (t[32, 1] - t0) / 1e3, "body sum", (t[:, 1] - t[:, 0]).sum() / 1e3)
