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


def test_c_base_placement_matches_numpy_oracle():
    """oracle/base_oracle.c (moment-matrix form, pthreads) against oracle/base_oracle.py (literal per-point residuals)."""
    import base_oracle as BO
    import c_oracle as CO
    from helpers import base_placement_case

    for name, B, n in (("panda_small", 4, 4), ("fetch_small", 3, 5), ("fetch10_c4", 2, 3)):
        t, qc, RTs, grid, origin, res = base_placement_case(name, B=B, n=n, spread=0.5)
        out = CO.base_place(t, qc, RTs, 0.01, grid, origin, res)
        one = CO.base_place(t, qc, RTs, 0.01, grid, origin, res, nthreads=1)
        assert np.array_equal(out["y"], one["y"]) and np.array_equal(out["Q"], one["Q"])  # threads do not change results
        for b in range(B):
            r = BO.solve_base(BO.BaseProblem(t, qc, RTs[b], 0.01, grid, origin, res))
            assert out["status"][b] == r.status and out["iters"][b] == r.iters
            tol = 1e-6 if r.status == 0 else 1e-4
            assert np.abs(out["y"][b] - r.y).max() < tol and np.abs(out["Q"][b] - r.Q).max() < tol
            assert abs(out["cost"][b] - r.cost) <= 1e-7 * max(1.0, r.cost)
            assert abs(out["collision"][b] - r.collision) <= (0 if np.abs(out["y"][b] - r.y).max() < 1e-9 else 2)
