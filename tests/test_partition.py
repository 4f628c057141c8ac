"""Graph partitioning with halo exchange (meshgraphnets.jl_b200/partition.py).
CPU half: the integer plan (ownership, halo lists, send/recv symmetry, edge coverage) - exact.
GPU half: P logical ranks on one device, stage-wise forward / backward with in-process exchanges, must
reproduce the unpartitioned step (outputs of owned rows, loss, summed parameter gradient)."""
import numpy as np
import pytest

import mgn_oracle as orc


def _mesh(nx, ny):
    pos, cells, nt = orc.cylinder_flow_mesh(nx, ny)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    return pos, nt, s, r


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_plan_is_exact(pkg, world):
    pos, nt, s, r = _mesh(11, 7)
    N, E = pos.shape[0], s.shape[0]
    parts = pkg.build_partition(N, s, r, world)
    assert [p.lo for p in parts] + [parts[-1].hi] == pkg.partition_bounds(N, world)
    # every edge belongs to exactly one rank: the owner of its receiver
    all_e = np.concatenate([p.edge_ids for p in parts])
    assert np.array_equal(np.sort(all_e), np.arange(E))
    for p in parts:
        glob = p.local_nodes_global()
        assert np.array_equal(glob[p.senders - 1], s[p.edge_ids] - 1)        # local ids map back to the global graph
        assert np.array_equal(glob[p.receivers - 1], r[p.edge_ids] - 1)
        assert ((p.receivers - 1) < p.n_own).all()                           # receivers are owned: scatter-sum is local
        assert np.array_equal(p.halo_global, np.unique(p.halo_global))
        assert not ((p.halo_global >= p.lo) & (p.halo_global < p.hi)).any()
        for q, rows in p.recv_rows.items():                                  # symmetric plan, same order on both sides
            sent = parts[q].send_rows[p.rank] + parts[q].lo
            assert np.array_equal(sent, p.halo_global[rows - p.n_own])
        one = pkg.build_partition_rank(N, s, r, world, p.rank)              # the single-rank builder agrees
        assert np.array_equal(one.senders, p.senders) and np.array_equal(one.receivers, p.receivers)
        assert np.array_equal(one.halo_global, p.halo_global) and np.array_equal(one.edge_ids, p.edge_ids)
        assert set(one.send_rows) == set(p.send_rows) and set(one.recv_rows) == set(p.recv_rows)
        assert all(np.array_equal(one.send_rows[q], p.send_rows[q]) for q in p.send_rows)
        assert all(np.array_equal(one.recv_rows[q], p.recv_rows[q]) for q in p.recv_rows)
        covered = np.sort(np.concatenate([v for v in p.recv_rows.values()])) if p.recv_rows else np.zeros(0, np.int64)
        assert np.array_equal(covered, np.arange(p.n_own, p.n_local))        # every halo row has exactly one owner


@pytest.mark.gpu
@pytest.mark.parametrize("mode,world,tol", [(0, 2, 2e-5), (0, 3, 2e-5), (1, 2, 2e-2), (1, 4, 2e-2)])
def test_partitioned_step_matches_unpartitioned(pkg, mode, world, tol):
    import torch
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rng = np.random.default_rng(5)
    pos, nt, s, r = _mesh(14, 9)
    N, E = pos.shape[0], s.shape[0]
    mps = 3
    model, ps, _ = pkg.build_model(9, 2, 2, mps, 128, 2, compute_mode=mode)
    nf = rng.normal(size=(N, 9)).astype(np.float32)
    ef = rng.normal(size=(E, 3)).astype(np.float32)
    tgt = rng.normal(size=(N, 2)).astype(np.float32)
    mask = orc.node_mask(nt, [0, 5])
    # ---- unpartitioned reference (same library, same mode)
    graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
    mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
    (g_ref,), loss_ref = pkg.step_(mgn, graph, dev(tgt), dev(mask))
    out_ref = model.forward(graph, ps).cpu().numpy()
    g_ref, loss_ref = g_ref.cpu().numpy(), float(loss_ref.cpu())
    # ---- P logical ranks on this device
    parts = pkg.build_partition(N, s, r, world)
    ranks, targets, masks = [], [], []
    for p in parts:
        glob = p.local_nodes_global()
        m = pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=mode)
        ranks.append(pkg.PartitionedModel(m, p, dev(nf[glob]), dev(ef[p.edge_ids])))
        targets.append(dev(tgt[glob]))
        own_mask = mask[(mask - 1 >= p.lo) & (mask - 1 < p.hi)] - p.lo       # 1-based local ids of owned masked nodes
        masks.append(dev(own_mask.astype(np.int32)))
    grads, losses, outs = pkg.run_partitioned_step(ranks, ps, targets, masks, len(mask), pkg.LocalExchange(),
                                                   pkg.masked_mse_partial)
    torch.cuda.synchronize()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    out = np.concatenate([o.cpu().numpy()[:p.n_own] for o, p in zip(outs, parts)])
    assert rel(out, out_ref) < tol
    loss = sum(float(l.cpu()) for l in losses)
    assert abs(loss - loss_ref) < tol * abs(loss_ref)
    g = sum(x.cpu().numpy().astype(np.float64) for x in grads)               # the gradient all-reduce (SUM)
    assert rel(g, g_ref) < (2e-4 if mode == 0 else 5e-2)
