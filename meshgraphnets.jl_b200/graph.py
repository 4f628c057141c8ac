"""Mirror of src/graph.jl: create_base_graph (:25-55) and build_graph (:75-97), over the C ABI."""
from __future__ import annotations

import numpy as np
import torch

import ctypes as C

from .core import (FeatureGraph, _ptr, _stream, call, edge_features, one_hot, parse_edges, shift_one_based,
                   triangles_to_edges)


def create_base_graph(data, type_size, type_min, device="cuda"):
    """src/graph.jl:25-55.  ``data`` maps names to arrays laid out [T, entities, features]
    (the C view of Julia's (features, entities, T)): "node_type" [1, N, 1], "mesh_pos" [1, N, dim],
    and "cells" [1, C, 3] or "edges" [U, 2].  Returns (node_type_onehot, senders, receivers,
    edge_features) on ``device`` - senders/receivers 1-based Int32 as in the reference.
    When the arrays are CUDA tensors (a trajectory already resident in HBM) everything runs on the device
    (create_base_graph_device) with bit-identical results."""
    if isinstance(data["mesh_pos"], torch.Tensor) and data["mesh_pos"].is_cuda:
        return create_base_graph_device(data, type_size, type_min)
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


def create_base_graph_device(data, type_size, type_min):
    """src/graph.jl:25-55 on the device: mgn_one_hot_device, mgn_triangles_to_edges_device (hash-set
    unique keeping the first occurrence) or mgn_parse_edges_device, mgn_shift_one_based_device and
    mgn_edge_features_device.  Inputs are CUDA tensors (node_type / cells / edges Int32, mesh_pos Float32)."""
    nt = data["node_type"][0].reshape(-1).to(torch.int32).contiguous()
    pos = data["mesh_pos"][0].to(torch.float32).contiguous()
    dev = pos.device
    depth = type_size - type_min + 1
    node_type = torch.empty((nt.shape[0], depth), dtype=torch.float32, device=dev)
    call("mgn_one_hot_device", _ptr(nt), nt.shape[0], int(depth), int(1 - type_min), _ptr(node_type), _stream())
    if "cells" in data:
        cells = data["cells"][0].reshape(-1, 3).to(torch.int32).contiguous()
        n = cells.shape[0]
        s = torch.empty(6 * n, dtype=torch.int32, device=dev)
        r = torch.empty(6 * n, dtype=torch.int32, device=dev)
        ne = C.c_int64(0)
        call("mgn_triangles_to_edges_device", _ptr(cells), n, _ptr(s), _ptr(r), C.byref(ne), _stream())
        E = ne.value
        U = E // 2
        # the kernel writes [hi ; lo] at [0, U) and [U, 2U) of buffers sized 6C: compact views of the first E entries
        senders, receivers = s[:E].clone(), r[:E].clone()
    elif "edges" in data:
        edges = data["edges"].reshape(-1, 2).to(torch.int32).contiguous()
        n = edges.shape[0]
        senders = torch.empty(2 * n, dtype=torch.int32, device=dev)
        receivers = torch.empty(2 * n, dtype=torch.int32, device=dev)
        call("mgn_parse_edges_device", _ptr(edges), n, _ptr(senders), _ptr(receivers), _stream())
    else:
        raise KeyError("Data does not contain cell or edge information!")
    flag = C.c_int32(0)
    call("mgn_shift_one_based_device", _ptr(senders), _ptr(receivers), senders.shape[0], C.byref(flag), _stream())
    ef = torch.empty((senders.shape[0], pos.shape[1] + 1), dtype=torch.float32, device=dev)
    call("mgn_edge_features_device", _ptr(pos), pos.shape[0], pos.shape[1], _ptr(senders), _ptr(receivers),
         senders.shape[0], 1, _ptr(ef), _stream())
    return node_type, senders, receivers, ef


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
