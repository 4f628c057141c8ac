"""Synthetic workloads of the BASELINE configs and the host-side index helpers of the training
driver.  Product code (numpy only): bench.py and the tools build their inputs here, never from oracle/."""
from __future__ import annotations

import numpy as np


def node_mask(node_type, types_updated, index_base=1):
    """src/MeshGraphNets.jl:352: Int32 ids (1-based) of the nodes whose type is in types_updated."""
    nt = np.asarray(node_type).reshape(-1)
    return (np.nonzero(np.isin(nt, list(types_updated)))[0] + index_base).astype(np.int32)


def val_mask(node_type, types_updated, out_dim):
    """src/MeshGraphNets.jl:354-358: Float32 0/1 mask repeated over the output rows -> [N, out]."""
    nt = np.asarray(node_type).reshape(-1)
    m = np.isin(nt, list(types_updated)).astype(np.float32)
    return np.repeat(m[:, None], out_dim, axis=1)


def cylinder_flow_mesh(nx=65, ny=29, lx=1.6, ly=0.41):
    """CylinderFlow-shaped structured triangulated grid: N = nx*ny = 1885 nodes, C = 2*(nx-1)*(ny-1) = 3584
    triangles (0-based Int32, same diagonal) -> 10 936 directed edges; node types x=0 -> 4, x=max -> 5,
    y=0 / y=max -> 6, interior 0, one interior column of type 1 (the inflow nodes of src/MeshGraphNets.jl:428)."""
    xs = np.linspace(0.0, lx, nx, dtype=np.float32)
    ys = np.linspace(0.0, ly, ny, dtype=np.float32)
    pos = np.stack(np.meshgrid(xs, ys, indexing="ij"), axis=-1).reshape(-1, 2).astype(np.float32)
    idx = np.arange(nx * ny, dtype=np.int32).reshape(nx, ny)
    a, b = idx[:-1, :-1].reshape(-1), idx[1:, :-1].reshape(-1)
    c, d = idx[1:, 1:].reshape(-1), idx[:-1, 1:].reshape(-1)
    cells = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], axis=0).astype(np.int32)
    nt = np.zeros((nx, ny), dtype=np.int32)
    nt[:, 0] = 6
    nt[:, -1] = 6
    nt[0, :] = 4
    nt[-1, :] = 5
    nt[1, 1:-1] = 1
    return pos, cells, nt.reshape(-1)


def synthetic_velocity(pos, T, seed=1234, noise=0.1):
    """Smooth travelling-wave field + N(0, noise^2), [T, N, 2] Float32 (numpy PCG64 stream)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(T, dtype=np.float32)[:, None] * 0.01
    x, y = pos[None, :, 0], pos[None, :, -1]
    u = np.sin(2 * np.pi * (x / 1.6 - t)) * np.cos(np.pi * y / 0.41) + 1.0
    v = 0.3 * np.cos(2 * np.pi * (x / 1.6 + t)) * np.sin(np.pi * y / 0.41)
    vel = np.stack([u, v], axis=-1)
    return (vel + rng.normal(0.0, noise, size=vel.shape)).astype(np.float32)


def chain_edges(n):
    """src/dataset.jl:379-382: 1-D chain, 1-based node pairs [i, i+1] (the input of parse_edges)."""
    i = np.arange(1, n, dtype=np.int32)
    return np.stack([i, i + 1], axis=1)


def tet_grid_edges(n):
    """Unique undirected edges of the Kuhn subdivision of an n^3 node grid, 1-based, lexicographically sorted
    [U, 2] (the format src/dataset.jl:345 hands to parse_edges): 3 axis, 3 face-diagonal and 1 body-diagonal
    neighbour per node -> 14 directed edges per interior node."""
    idx = np.arange(n ** 3, dtype=np.int64).reshape(n, n, n)
    out = []
    for dx, dy, dz in ((0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)):
        a = idx[:n - dx, :n - dy, :n - dz].reshape(-1)
        b = idx[dx:, dy:, dz:].reshape(-1)
        out.append(np.stack([a, b], 1))
    e = np.concatenate(out)
    e = e[np.lexsort((e[:, 1], e[:, 0]))]
    return (e + 1).astype(np.int32)
