"""Generates the golden vectors of tests/golden/ from the CPU oracle (oracle/mgn_oracle.py, fp64).
The reference itself cannot run here (no Julia / GraphNetCore.jl - SURVEY.md 8c), so these pin the
ORACLE, not the reference: `oracle/julia/dump_reference.jl` replays the same inputs through real
GraphNetCore for a maintainer who has it.    python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import mgn_oracle as orc  # noqa: E402
import mgn_oracle_bf16 as ob  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    # ---- integer path: a 6 x 4 triangulated grid and a 9-node chain
    pos, cells, nt = orc.cylinder_flow_mesh(6, 4)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    rp, perm = orc.build_csr(r, pos.shape[0])
    cp, perm_s = orc.build_csr(s, pos.shape[0])
    es, er = orc.parse_edges(orc.create_edges_1d(9))
    crp, cperm = orc.build_csr(er, 9)
    np.savez(os.path.join(HERE, "index_golden.npz"), pos=pos, cells=cells, node_type=nt, senders=s, receivers=r,
             row_ptr=rp, perm=perm, col_ptr=cp, perm_sender=perm_s, onehot=orc.one_hot(nt, 7, 1),
             edge_features=orc.edge_features(pos, s, r), mask=orc.node_mask(nt, [0, 5]),
             chain_senders=es, chain_receivers=er, chain_row_ptr=crp, chain_perm=cperm)
    # ---- float path: 2 MP steps, latent 128, on the same grid
    cfg = orc.ModelConfig(9, 3, 2, 128, 2, 2)
    ps = (orc.init_params(cfg, seed=7, dtype=np.float64) + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    N, E = pos.shape[0], s.shape[0]
    nf = rng.normal(size=(N, 9)).astype(np.float32)
    ef = rng.normal(size=(E, 3)).astype(np.float32)
    tgt = rng.normal(size=(N, 2)).astype(np.float32)
    mask = orc.node_mask(nt, [0, 5])
    g, loss, out, dnf = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    np.savez(os.path.join(HERE, "cyl_small_inputs.npz"), node_in=9, edge_in=3, out_dim=2, latent=128, mps=2,
             hidden_layers=2, params=ps, nf=nf, ef=ef, senders=s, receivers=r, target=tgt, mask=mask)
    # gradients are stored as per-tensor norms + a strided sample to keep the fixture small
    specs, P = orc.mlp_specs(cfg)
    # the same step in the tensor-core mode's arithmetic (bf16 roundings where the kernels store bf16)
    g_b, loss_b, out_b, _ = ob.step_bf16(cfg, ps, nf, ef, s, r, tgt, mask)
    np.savez(os.path.join(HERE, "cyl_small_golden.npz"), out=out, loss=loss, dnf=dnf,
             out_bf16=out_b, loss_bf16=loss_b, grad_sample_bf16=g_b[::97].copy(), grad_norm_bf16=np.linalg.norm(g_b),
             grad_sample=g[::97].copy(), grad_norm=np.linalg.norm(g),
             tensor_norms=np.array([np.linalg.norm(g[sp.offset:sp.offset + sp.size]) for sp in specs]))
    # ---- Adam and online normaliser
    p = rng.normal(size=257).astype(np.float32)
    m = np.zeros_like(p); v = np.zeros_like(p)
    gs = rng.normal(size=(3, 257)).astype(np.float32)
    traj = [p.copy()]
    for t in range(1, 4):
        p, m, v = orc.adam_update(p, gs[t - 1], m, v, t, lr=1e-4)
        traj.append(p.copy())
    on = orc.NormaliserOnline(3)
    xs = (rng.normal(size=(2, 50, 3)) * [1, 10, 0.1] + [0, 5, -2]).astype(np.float32)
    ys = np.stack([on(x) for x in xs])
    np.savez(os.path.join(HERE, "optim_norm_golden.npz"), adam_grads=gs, adam_traj=np.stack(traj), norm_x=xs,
             norm_y=ys, norm_sum=on.acc_sum, norm_sum_sq=on.acc_sum_sq, norm_count=on.acc_count)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
