"""Per-role timestamp traces (CTA 0) of the tensor-core kernels inside a bench-size training step.
Usage: python tools/trace_kernels.py [batch]   -> prints, per kernel family, the event series in us."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import mgn_oracle as orc  # noqa: E402
import mgn_pkg  # noqa: E402

pkg = mgn_pkg.pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pos, cells, nt = orc.cylinder_flow_mesh(65, 29)
N0 = pos.shape[0]
cells_b = np.concatenate([cells + b * N0 for b in range(B)], axis=0)
s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells_b))
N, E = N0 * B, s.shape[0]
rng = np.random.default_rng(0)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
nf = dev(rng.normal(size=(N, 9)).astype(np.float32))
ef = dev(rng.normal(size=(E, 3)).astype(np.float32))
tgt = dev(rng.normal(size=(N, 2)).astype(np.float32))
mask = dev(orc.node_mask(np.tile(nt, B), [0, 5]))
model, ps, _ = pkg.build_model(9, 2, 2, 15, 128, 2, compute_mode=pkg.COMPUTE_BF16)
graph = pkg.FeatureGraph(nf, ef, dev(s), dev(r))
mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
lib = pkg.load()   # needs the debug build: python meshgraphnets.jl_b200/build.py --trace
lib.mgn_debug_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
for _ in range(2):
    pkg.step_(mgn, graph, tgt, mask)
torch.cuda.synchronize()

ROLES = {0: "head", 1: "mma", 2: "epi", 3: "prod"}
# launch order inside a step: fwd: enc.node, enc.edge, (edge, node) x 15, dec ; bwd chain: dec, (node, edge) x 15, ...
for fam, skip, name in ((0, 2 + 2 * 7, "forward, edge MLP of MP step 8"), (1, 1 + 2 * 7 + 1, "bwd chain, edge MLP of MP step 8"),
                        (2, 1 + 2 * 7 + 1, "bwd input, edge MLP of MP step 8")):
    buf = torch.zeros(4 * 512, dtype=torch.int64, device="cuda")
    lib.mgn_debug_trace(buf.data_ptr(), fam, skip)
    pkg.step_(mgn, graph, tgt, mask)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(4, 512)
    t0 = t[t > 0].min() if (t > 0).any() else 0
    print(f"==== {name}: E={E} N={N}")
    for role in range(4):
        ev = t[role][t[role] > 0]
        if len(ev) == 0:
            continue
        us = (ev - t0) / 1e3
        print(f"  {ROLES[role]:5s} n={len(ev):3d}: " + " ".join(f"{x:.1f}" for x in us[:90]))
