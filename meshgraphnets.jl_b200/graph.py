"""Mirror of src/graph.jl: create_base_graph (:25-55) and build_graph (:75-97), over the C ABI."""
from __future__ import annotations

import numpy as np
import torch

from .core import (FeatureGraph, edge_features, one_hot, parse_edges, shift_one_based,
                   triangles_to_edges)


def create_base_graph(data, type_size, type_min, device="cuda"):
    """src/graph.jl:25-55.  ``data`` maps names to arrays laid out [T, entities, features]
    (the C view of Julia's (features, entities, T)): "node_type" [1, N, 1], "mesh_pos" [1, N, dim],
    and "cells" [1, C, 3] or "edges" [U, 2].  Returns (node_type_onehot, senders, receivers,
    edge_features) on ``device`` - senders/receivers 1-based Int32 as in the reference."""
    node_type = one_hot(np.asarray(data["node_type"])[0].reshape(-1), type_size - type_min + 1, 1 - type_min)
    if "cells" in data:
        senders, receivers = triangles_to_edges(np.asarray(data["cells"])[0])
    elif "edges" in data:
        senders, receivers = parse_edges(np.asarray(data["edges"]))
    else:
        raise KeyError("Data does not contain cell or edge information!")
    shift_one_based(senders, receivers)
    ef = edge_features(np.asarray(data["mesh_pos"])[0], senders, receivers, index_base=1)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return to(node_type), to(senders), to(receivers), to(ef)


def build_graph(mgn, data, fields, datapoint, node_type, edge_feats, senders, receivers):
    """src/graph.jl:75-97.  nf = vcat(n_norm[f](data[f][:, :, min(T, datapoint)]) for f in fields...,
    n_norm["node_type"](node_type)); ef = e_norm(edge_features).  The normalisers write straight
    into the concatenated matrix (no vcat reallocations).  ``datapoint`` is 1-based."""
    N = node_type.shape[0]
    widths = [data[f].shape[-1] for f in fields]
    nf = torch.empty((N, sum(widths) + node_type.shape[1]), dtype=torch.float32, device=node_type.device)
    # the reference evaluates the node_type normaliser first (graph.jl:80); it lands last (graph.jl:86)
    mgn.n_norm["node_type"](node_type, out=nf, col=sum(widths))
    col = 0
    for f, w in zip(fields, widths):
        x = data[f]
        t = min(x.shape[0], datapoint) - 1 if x.dim() == 3 else None
        mgn.n_norm[f](x[t] if t is not None else x, out=nf, col=col)
        col += w
    ef = mgn.e_norm(edge_feats)
    return FeatureGraph(nf, ef, senders, receivers)
