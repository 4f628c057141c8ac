"""Compares the output of oracle/julia/dump_reference.jl (real GraphNetCore.jl, run by a maintainer who has Julia)
with the oracle's golden vectors in this directory and names, for every mismatch, the recalled semantic of DESIGN.md
section 5 that has to be flipped.  Test infrastructure; nothing imports it.

    python tests/golden/compare_reference.py out_dir/reference_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# key -> (golden file, golden key, exact?, what a mismatch means)
CHECKS = {
    "senders": ("index_golden.npz", "senders", True, "triangles_to_edges: edge order / (max, min) convention (SURVEY 9.7)"),
    "receivers": ("index_golden.npz", "receivers", True, "triangles_to_edges: edge order / (max, min) convention (SURVEY 9.7)"),
    "onehot": ("index_golden.npz", "onehot", True, "one_hot(v, depth, offset) row convention"),
    "chain_senders": ("index_golden.npz", "chain_senders", True, "parse_edges two-way order (SURVEY 9.7)"),
    "chain_receivers": ("index_golden.npz", "chain_receivers", True, "parse_edges two-way order (SURVEY 9.7)"),
    "edge_features": ("index_golden.npz", "edge_features", True, "norm accumulation width of the edge length (SURVEY 9.9)"),
    "out": ("cyl_small_golden.npz", "out", False,
            "model forward: MLP depth, LayerNorm placement / epsilon / (bias, scale) order, concat order, pre-residual "
            "aggregation, parameter flattening order (SURVEY 9.1-9.4, 9.10)"),
    "loss": ("cyl_small_golden.npz", "loss", False, "step! loss: mse_reduce row sum, mean over masked nodes (SURVEY 9.6)"),
    "grad_norm": ("cyl_small_golden.npz", "grad_norm", False, "gradient of step! (follows from the forward semantics)"),
    "grad_sample": ("cyl_small_golden.npz", "grad_sample", False, "gradient layout: parameter flattening order (SURVEY 9.4)"),
    "adam_traj": ("optim_norm_golden.npz", "adam_traj", False, "Optimisers.Adam epsilon / bias-correction order (SURVEY 9.8)"),
    "norm_y": ("optim_norm_golden.npz", "norm_y", False, "NormaliserOnline accumulate-then-normalise, std_epsilon (SURVEY 9.5)"),
}


def main(path):
    ref = np.load(path, allow_pickle=True)
    bad = 0
    for key, (fname, gkey, exact, meaning) in CHECKS.items():
        if key not in ref:
            print(f"MISSING  {key}")
            bad += 1
            continue
        want = np.load(os.path.join(HERE, fname))[gkey]
        got = np.asarray(ref[key])
        if got.shape != want.shape and got.ndim == 3 and want.ndim == 3:      # (rows, F, calls) vs (calls, rows, F)
            got = np.moveaxis(got, -1, 0)
        if got.shape != want.shape:
            print(f"SHAPE    {key}: reference {got.shape} vs oracle {want.shape}  -> {meaning}")
            bad += 1
            continue
        if exact:
            ok = np.array_equal(got, want)
            err = float(np.abs(got.astype(np.float64) - want).max())
        else:
            err = float(np.linalg.norm(got.astype(np.float64) - want) / max(np.linalg.norm(want), 1e-30))
            ok = err < 1e-4
        print(f"{'ok      ' if ok else 'MISMATCH'} {key}: err {err:.3e}" + ("" if ok else f"  -> {meaning}"))
        bad += not ok
    if "param_labels" in ref:
        print("ComponentArray labels of the reference (compare with mgn_model_param_layout):")
        print(str(ref["param_labels"])[:2000])
    print("PINNED: the oracle reproduces GraphNetCore on these inputs" if bad == 0 else f"{bad} check(s) failed")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(sys.argv[1]) else 0)
