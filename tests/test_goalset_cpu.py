"""Goal-set selection (grasptrajopt_b200/goalset.py): the reference's ``mmin`` over per-goal costs (gto/gto_planner.py:91-105)
as an arg-min after the batch solve, ranked by status class so that a NaN solve (reported cost 0 before the fix) or a
non-converged goal never beats a converged one."""
import numpy as np

from grasptrajopt_b200.goalset import goalset_best, status_class


def test_converged_goals_win_over_cheaper_unconverged_ones():
    cost = np.array([0.5, 0.1, 0.3, 0.0])
    status = np.array([0, 1, 0, 2])  # converged, max_iter (cheaper), converged, NaN with a fake cost of 0
    assert goalset_best(cost, status) == 2
    assert status_class(cost, status).tolist() == [0, 1, 0, 2]


def test_fallbacks_and_ties():
    assert goalset_best([0.3, 0.2, 0.2], [1, 4, 3]) == 1  # nobody converged: cheapest finite cost, lowest index on ties
    assert goalset_best([np.nan, np.inf], [2, 2]) == 0  # everything failed: still returns an index
    assert goalset_best([0.2, np.inf, 0.1], [0, 0, 2]) == 0  # non-finite cost counts as failed whatever the status says
    assert goalset_best([3.0, 1.0, 2.0]) == 1  # no status: plain arg-min
