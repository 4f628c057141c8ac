"""Roofline of the dominant kernel + launch count, measured live by bench.py (rank 0).

A few extra, un-captured steps run with the library's launch profiler on: every launch of the
dominant kernel family is bracketed by CUDA events on its own stream, inside the real step.
achieved = algorithmic FLOPs (or bytes) of those launches / their summed device time.
Algorithmic work per unit follows SURVEY.md 8d / DESIGN.md:
  edge MLP forward : 2*D^2*(L+2) FLOP per edge          (L = hidden_layers + 2 Dense layers)
  node MLP forward : 2*D^2*(L+1) FLOP per node
  backward         : 2x forward (dX and dW GEMMs; recompute is NOT counted)
"""
from __future__ import annotations

import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def roofline_and_launches(args, pkg, model, mgn, E, B, dev, step_fn=None, n_nodes=None):
    from bench import HIDDEN, LATENT, MPS
    D, L = LATENT, HIDDEN + 2
    out = {}
    if step_fn is None:
        return out
    # launches per step
    pkg.profile_begin(-1)
    step_fn()
    n_launch, _, _, per = pkg.profile_end()
    out["gpu_launches_per_step"] = int(n_launch)
    out["launches_by_kernel"] = per
    pk = peaks()
    fams = []
    if args.mode == "bf16":
        fams = [("tc_mlp_fwd", 2.0 * D * D * ((L + 2) * E + (L + 1) * n_nodes) * MPS)]
    else:
        fams = [("simt_gemm_fwd", 2.0 * D * D * ((L + 2) * E + (L + 1) * n_nodes) * MPS)]
    for name, flops in fams:
        if name not in per:
            continue
        reps = 3
        pkg.profile_begin(pkg.profile_tag(name))
        for _ in range(reps):
            step_fn()
        _, k, ms, _ = pkg.profile_end()
        if k == 0 or ms <= 0:
            continue
        tf = flops * reps / (ms * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        out["roofline"] = {"kernel": name, "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                           "frac": tf / peak, "traffic": None, "launches_timed": int(k),
                           "avg_launch_us": 1e3 * ms / k,
                           "note": f"algorithmic fwd FLOPs of the processor blocks (encoder/decoder launches of the "
                                   f"same family are timed but their FLOPs not counted); peak = "
                                   f"{pk['source']} sustained bf16 (kernel timed inside a long step)"}
    return out
