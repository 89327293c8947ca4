"""The C restatement of the oracle (CPU baseline) against the NumPy oracle.  CPU only."""
import numpy as np

import c_oracle
import gto_oracle as O
from helpers import problems_from_workload, small_workload


def test_c_oracle_matches_numpy_oracle():
    w = small_workload("C2", "panda_small", B=4, n_field=48)
    rc = c_oracle.solve_workload(w, nthreads=2)
    for i, p in enumerate(problems_from_workload(w)):
        r = O.solve_lm(p)
        assert rc["status"][i] == r.status and rc["iters"][i] == r.iters
        np.testing.assert_allclose(rc["Q"][i], r.Q, atol=1e-9)
        np.testing.assert_allclose(rc["dQ"][i], r.dQ, atol=1e-9)
        assert abs(rc["cost"][i] - r.cost) < 1e-10


def test_c_oracle_fetch_prismatic_and_zero_field():
    w = small_workload("C3", None, B=2, n_field=48)
    w.batch.field_all[:] = -1
    rc = c_oracle.solve_workload(w, nthreads=2)
    for i, p in enumerate(problems_from_workload(w)):
        r = O.solve_lm(p)
        assert rc["iters"][i] == r.iters
        np.testing.assert_allclose(rc["Q"][i], r.Q, atol=1e-8)


def test_c_oracle_threads_do_not_change_results():
    w = small_workload("C2", "panda_small", B=5, n_field=48)
    a = c_oracle.solve_workload(w, nthreads=1)
    b = c_oracle.solve_workload(w, nthreads=4)
    np.testing.assert_array_equal(a["Q"], b["Q"])
    np.testing.assert_array_equal(a["iters"], b["iters"])
