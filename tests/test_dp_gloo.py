"""N > 1 host logic on CPU: world_size-2 gloo run of the data-parallel step semantics
(meshgraphnets.jl_b200/parallel.py).  The DP gradient must equal the mean of the single-window oracle
gradients (batch-P SGD), windows must be sharded disjointly, and the online-normaliser
statistics must equal those of a serial pass over all windows."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import mgn_oracle as orc
    pos, cells, nt = orc.cylinder_flow_mesh(5, 4)
    s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
    cfg = orc.ModelConfig(4, 3, 2, 16, 2, 1)
    ps = orc.init_params(cfg, seed=5, dtype=np.float64)
    rng = np.random.default_rng(11)
    N, E = pos.shape[0], s.shape[0]
    windows = [(rng.normal(size=(N, 4)), rng.normal(size=(E, 3)), rng.normal(size=(N, 2))) for _ in range(4)]
    return orc, cfg, ps, s, r, orc.node_mask(nt, [0, 5]), windows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc, cfg, ps, s, r, mask, windows = _problem()
        import mgn_pkg
        par = mgn_pkg.pkg
        mine = par.shard_windows(len(windows), rank, world)
        g = np.zeros_like(ps)
        norm = orc.NormaliserOnline(4)
        prev = np.concatenate([norm.acc_sum, norm.acc_sum_sq, [norm.acc_count, norm.num_acc]]).astype(np.float32)
        for w in mine:                               # one window per rank per step; steps accumulate here
            nf, ef, tgt = windows[w]
            gw, *_ = orc.step(cfg, ps, nf, ef, s, r, tgt, mask, dtype=np.float64)
            g += gw
            norm(nf.astype(np.float32))
        g /= len(mine)
        flat = torch.from_numpy(g.copy())
        par.allreduce_mean_(flat, world)
        state = torch.from_numpy(np.concatenate([norm.acc_sum, norm.acc_sum_sq,
                                                 [norm.acc_count, norm.num_acc]]).astype(np.float32))
        par.allreduce_normaliser_(state, torch.from_numpy(prev))
        if rank == 0:
            q.put((mine, flat.numpy(), state.numpy()))
        else:
            q.put((mine, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp2_gradient_is_mean_of_window_gradients():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = sorted(m for m, _, _ in res)
    assert shards == [[0, 2], [1, 3]]                      # disjoint, covering, strided
    flat, state = next((f, s) for _, f, s in res if f is not None)
    orc, cfg, ps, s, r, mask, windows = _problem()
    ref = np.mean([orc.step(cfg, ps, nf, ef, s, r, tgt, mask, dtype=np.float64)[0] for nf, ef, tgt in windows], axis=0)
    assert np.linalg.norm(flat - ref) < 1e-12 * np.linalg.norm(ref)
    serial = orc.NormaliserOnline(4)
    for nf, _, _ in windows:
        serial(nf.astype(np.float32))
    assert np.allclose(state[:4], serial.acc_sum, rtol=1e-5) and state[8] == serial.acc_count
    assert state[9] == serial.num_acc


def _shooting_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
        import mgn_pkg
        from test_oracle_solver import _problem as shooting_problem
        from test_shooting_host import _run
        pkg = mgn_pkg.pkg
        cfg, p, make, _, N = shooting_problem(seed=3, T=9)
        kw = dict(tstart=0.0, dt=0.01, tstop=0.08, interval_size=3, continuity_term=10, solver="euler", n_sub=1)
        owned = pkg.shard_intervals(len(pkg.shooting_ranges(9, 3)), rank, world)
        g, loss, _, _ = _run(pkg, make, p, N, owned=owned, **kw)
        tg, tl = torch.from_numpy(g.copy()), torch.tensor([loss], dtype=torch.float64)
        pkg.allreduce_sum_(tg, tl)
        q.put((owned, tg.numpy() if rank == 0 else None, float(tl[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_multiple_shooting_intervals_sharded_over_2_ranks():
    """shooting intervals are independent solves; sharded over two gloo ranks and SUM-reduced, loss and
    gradient equal the unsharded step (the engine runs on the CPU test doubles of tests/test_shooting_host.py)."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shooting_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(o for o, _, _ in res) == [[0, 2], [1, 3]]
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import mgn_oracle_solver as sol
    from test_oracle_solver import _problem as shooting_problem
    cfg, p, make, _, N = shooting_problem(seed=3, T=9)
    g_ref, loss_ref, _ = sol.train_step_multiple_shooting(make(p), 0.0, 0.01, 0.08, 3, 10, "euler", 1)
    g = next(f for _, f, _ in res if f is not None)
    assert all(abs(l - loss_ref) < 1e-11 * abs(loss_ref) for _, _, l in res)
    assert np.allclose(g, g_ref, rtol=1e-9, atol=1e-12 * np.abs(g_ref).max())


def test_shard_windows_properties():
    sys.path[:0] = [ROOT]
    import mgn_pkg
    for n in (0, 1, 7, 600):
        for world in (1, 2, 4, 8):
            parts = [mgn_pkg.pkg.shard_windows(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
