"""Mirror of the training-strategy entry points of src/strategies.jl that reach the hot path."""
from __future__ import annotations

import torch

from .core import mse_reduce, step_
from .graph import build_graph


class DerivativeTraining:
    """src/strategies.jl:389-447."""

    def __init__(self, window_size=0, random=True):
        self.window_size, self.random = window_size, random


def get_delta(strategy, trajectory_length):
    """src/strategies.jl:391-393."""
    return strategy.window_size if strategy.window_size > 0 else trajectory_length - 1


def init_train_step(strategy, t, ta=None):
    """src/strategies.jl:395-415: target = o_norm[f]((data["target|f"][t] - data[f][t]) / dt) for
    each target field (the online normaliser accumulates here), then build_graph."""
    mgn, data, meta, fields, target_fields, node_type, edge_feats, senders, receivers, datapoint, mask, _ = t
    cols = []
    for f in target_fields:
        cur, nxt = data[f][datapoint - 1], data["target|" + f][datapoint - 1]
        if isinstance(meta["dt"], (list, tuple)):
            dt = meta["dt"][datapoint] - meta["dt"][datapoint - 1]
        else:
            dt = float(meta["dt"])
        cols.append(mgn.o_norm[f]((nxt - cur) / dt))
    target = torch.cat(cols, dim=1) if len(cols) > 1 else cols[0]
    graph = build_graph(mgn, data, fields, datapoint, node_type, edge_feats, senders, receivers)
    return mgn, graph, target, mask


def train_step(strategy, t):
    """src/strategies.jl:417-422."""
    mgn, graph, target, mask = t
    return step_(mgn, graph, target, mask, mse_reduce)
