"""Goal-set selection shared by the single-GPU planner shim and the multi-GPU reduction.

The reference couples the goals of ``plan_goalset`` through ``mmin`` over per-goal costs (``gto/gto_planner.py:91-105``); here
the goals are independent problems of one batch and the winner is picked afterwards.  A cost is only comparable when the
solve behind it produced a usable trajectory, so candidates are ranked by status class first:

    class 0   converged (``STATUS_CONVERGED``)
    class 1   finite cost, not converged (``MAX_ITER``, ``SLOW``: rests on a kink / stopped early, ``STALLED``)
    class 2   ``STATUS_NAN`` or a non-finite cost -- never wins unless every candidate is in this class

then by cost, then by index (lowest wins ties).
"""
from __future__ import annotations

import numpy as np

STATUS_CONVERGED, STATUS_NAN = 0, 2


def status_class(cost, status) -> np.ndarray:
    cost = np.asarray(cost, dtype=np.float64)
    status = np.asarray(status)
    cls = np.where(status == STATUS_CONVERGED, 0, 1)
    cls = np.where((status == STATUS_NAN) | ~np.isfinite(cost), 2, cls)
    return cls.astype(np.int64)


def goalset_best(cost, status=None) -> int:
    """Index of the winning goal (see module docstring).  ``status`` None: every candidate counts as converged."""
    cost = np.asarray(cost, dtype=np.float64).reshape(-1)
    if cost.size == 0:
        raise ValueError("empty goal set")
    cls = status_class(cost, np.zeros(cost.size, np.int64) if status is None else np.asarray(status).reshape(-1))
    key = np.where(np.isfinite(cost), cost, np.inf)
    return int(np.lexsort((np.arange(cost.size), key, cls))[0])
