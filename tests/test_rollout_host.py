"""Host logic of the rollout mirror (src/solve.jl:42-68): the fixed-step grid that
``solve(prob, solver; adaptive=false, dt=dt, saveat=saves)`` walks over (start, stop).  No GPU needed."""
import numpy as np
import pytest


def _plan(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".solve")._step_plan


def test_one_step_per_save_when_dt_equals_the_spacing(pkg):
    saves = [np.float32(0.01) * i for i in range(51)]
    start, ts, plan = _plan(pkg)(0.0, 0.5, 0.01, saves)
    assert start == 0.0 and len(ts) == 51 and len(plan) == 51
    assert plan[0][0] == 0                       # saves[0] == start: the initial state is the first save
    assert all(n == 1 for n, _, _ in plan[1:]) and all(h == np.float32(0.01) for _, h, _ in plan[1:])
    assert [t0 for _, _, t0 in plan[1:]] == [np.float32(v) for v in saves[:-1]]      # interval starts are the save times


def test_sub_steps_when_dt_divides_the_spacing(pkg):
    """dt = 0.005 with saves every 0.01: two steps per save (the round-1 mirror took one and mislabelled the time)."""
    start, ts, plan = _plan(pkg)(0.0, 0.1, 0.005, [0.0, 0.01, 0.02, 0.04])
    assert [n for n, _, _ in plan] == [0, 2, 2, 4]


def test_lead_in_from_start_to_the_first_save(pkg):
    start, ts, plan = _plan(pkg)(0.0, 0.1, 0.01, [0.02, 0.03])
    assert [n for n, _, _ in plan] == [2, 1] and start == 0.0


def test_dt_none_is_the_tstops_branch(pkg):
    start, ts, plan = _plan(pkg)(0.0, 0.1, None, [0.0, 0.01, 0.03])
    assert [n for n, _, _ in plan] == [0, 1, 1]
    assert abs(float(plan[2][1]) - 0.02) < 1e-7


@pytest.mark.parametrize("args", [
    (0.0, 0.1, 0.004, [0.0, 0.01]),            # dt does not divide the spacing: would need dense output
    (0.0, 0.05, 0.01, [0.0, 0.06]),            # save beyond stop
    (0.02, 0.1, 0.01, [0.0, 0.01]),            # save before start
    (0.0, 0.1, 0.01, [0.0]),                   # saves[2] - saves[1] is undefined
])
def test_grids_the_mirror_cannot_honour_raise(pkg, args):
    with pytest.raises(pkg.MgnError):
        _plan(pkg)(*args)
