"""CPU checks of oracle/mgn_oracle_bf16.py (the model of the tensor-core mode's arithmetic): the
bf16 rounding helper against known answers, and the distance between bf16-operand arithmetic and
the fp64 restatement of the reference - the number DESIGN.md quotes as the bf16-mode tolerance."""
import numpy as np

import mgn_oracle as orc
import mgn_oracle_bf16 as ob


def test_bf16_rounding_known_answers():
    x = np.array([1.0, 1.00390625, 1.001953125, 1.005859375, -2.5, 3.3895313892515355e38, 0.0], np.float32)
    # 1 + 2^-8 is a tie between 1.0 and 1 + 2^-7 -> even (1.0); 1 + 3*2^-9... etc.
    got = ob.q(x)
    assert got[0] == 1.0 and got[1] == 1.0 and got[2] == 1.0
    assert got[3] == 1.0078125                      # 1 + 1.5 * 2^-8 rounds up to 1 + 2^-7
    assert got[4] == -2.5 and got[6] == 0.0
    import torch
    r = torch.randn(10000, dtype=torch.float32)
    assert np.array_equal(ob.q(r.numpy()), r.bfloat16().double().numpy())


def _problem(nx, ny, mps, hidden=2, seed=0):
    rng = np.random.default_rng(seed)
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, mps, hidden)
    ps = (orc.init_params(cfg, seed=seed + 1, dtype=np.float64)
          + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    nf = rng.normal(size=(N, 9)).astype(np.float32)
    ef = rng.normal(size=(E, 3)).astype(np.float32)
    tgt = rng.normal(size=(N, 2)).astype(np.float32)
    return cfg, ps, nf, ef, s, r, tgt, orc.node_mask(nt, [0, 5])


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def test_bf16_model_is_close_to_fp64_oracle():
    """The cost of bf16 operands on a 3-step model: loss within 1e-2, gradients within 6e-2."""
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(12, 9, 3)
    g_o, loss_o, out_o, dnf_o = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    g_b, loss_b, out_b, dnf_b = ob.step_bf16(cfg, ps, nf, ef, s, r, tgt, mask)
    assert abs(loss_b - loss_o) < 1e-2 * abs(loss_o)
    assert rel(out_b, out_o) < 3e-2
    assert rel(g_b, g_o) < 6e-2
    assert rel(dnf_b, dnf_o) < 0.2


def test_bf16_model_without_rounding_is_the_oracle(monkeypatch):
    """With q() replaced by the identity the model must reproduce the fp64 oracle (structure check
    of the restated backward: CSR/CSC orders, sinks, decoder head)."""
    monkeypatch.setattr(ob, "q", lambda x: np.asarray(x, np.float64))
    monkeypatch.setattr(ob, "f32", lambda x: np.asarray(x, np.float64))
    cfg, ps, nf, ef, s, r, tgt, mask = _problem(7, 5, 2, hidden=1)
    g_o, loss_o, out_o, dnf_o = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    g_b, loss_b, out_b, dnf_b = ob.step_bf16(cfg, ps.astype(np.float64), nf.astype(np.float64), ef.astype(np.float64),
                                             s, r, tgt, mask)
    assert rel(out_b, out_o) < 1e-5 and rel(g_b, g_o) < 1e-5 and rel(dnf_b, dnf_o) < 1e-5
