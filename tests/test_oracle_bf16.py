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


def test_input_gradient_error_comes_from_the_forward_point_not_from_the_vjp_arithmetic():
    """Why the d/d(node features) tolerance of the tensor-core mode is loose (0.2-0.3) while the parameter gradient's is
    tight: the input Jacobian of a deep ReLU message-passing network is DISCONTINUOUS in the
    forward point (gate flips), so any perturbation of the forward activations - here the bf16 operand rounding -
    moves it by O(10 %) per node, while the parameter gradient averages over all rows.  Pinned with the CPU model of
    the kernels' arithmetic: (a) with the bf16 forward kept, making the WHOLE backward exact leaves the d/d nf error
    unchanged; (b) with an exact forward and the bf16 backward the error collapses by an order of magnitude.  So fp32
    accumulation of the encoder-input VJP (or any other change of the backward arithmetic) cannot tighten it; only a
    higher-precision forward (the fp32 mode: 1e-5) does."""
    import mgn_oracle as orc
    import mgn_oracle_bf16 as ob
    rng = np.random.default_rng(0)
    pos, cells, nt = orc.cylinder_flow_mesh(24, 14)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    N, E = pos.shape[0], s.shape[0]
    cfg = orc.ModelConfig(9, 3, 2, 128, 8, 2)
    ps = (orc.init_params(cfg, seed=1, dtype=np.float64) + 0.02 * rng.normal(size=orc.mlp_specs(cfg)[1])).astype(np.float32)
    nf, ef = rng.normal(size=(N, 9)).astype(np.float32), rng.normal(size=(E, 3)).astype(np.float32)
    tgt, mask = rng.normal(size=(N, 2)).astype(np.float32), orc.node_mask(nt, [0, 5])
    g64, _, _, d64 = orc.step(cfg, ps.astype(np.float64), nf, ef, s, r, tgt, mask, dtype=np.float64)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    q0, lad0 = ob.q, orc.loss_and_dout
    state = {"bwd": False}

    def lad(*a, **k):                       # the loss is the boundary between the forward and the backward sweep
        state["bwd"] = True
        return lad0(*a, **k)

    def run(fwd_q, bwd_q):
        state["bwd"] = False
        ob.q = lambda x: q0(x) if (bwd_q if state["bwd"] else fwd_q) else np.asarray(x, np.float64)
        orc.loss_and_dout = lad
        try:
            g, _, _, d = ob.step_bf16(cfg, ps, nf, ef, s, r, tgt, mask)
        finally:
            ob.q, orc.loss_and_dout = q0, lad0
        return rel(g, g64), rel(d, d64)

    g_all, d_all = run(True, True)
    g_fq, d_fq = run(True, False)           # bf16 forward, exact backward
    g_bq, d_bq = run(False, True)           # exact forward, bf16 backward
    assert d_all > 0.03                                       # the effect is there at 8 MP steps already (0.19 at 15)
    assert abs(d_fq - d_all) < 0.15 * d_all                   # (a) the backward arithmetic does not matter
    assert d_bq < 0.2 * d_all                                 # (b) the forward point does
    assert g_all < 0.3 * d_all                                # the parameter gradient averages the flips out
