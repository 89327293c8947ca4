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


def test_active_set_round_and_status_semantics_match_between_c_and_numpy():
    """Fetch-8 shelf problems push joints against their limits: the active-set round (held variables get the step to the
    limit, the rest is re-solved) must engage, give the same iterates in both oracles, and beat plain clipping."""
    w = small_workload("C3", None, B=3, n_field=48)
    with_as = c_oracle.solve_workload(w, nthreads=2)
    clip = c_oracle.solve_workload(w, nthreads=2, options=c_oracle.default_options(as_rounds=0))
    assert np.any(with_as["iters"] != clip["iters"]) or np.abs(with_as["Q"] - clip["Q"]).max() > 1e-9  # the round changes the path
    assert with_as["cost"].sum() <= clip["cost"].sum() * (1 + 1e-6)
    # the two restatements walk the same path: identical iterates over the first 40 steps (bundle step, active-set round,
    # damping policy).  Later a problem that ends on kinks of the trilinear field takes accept / reject decisions on cost
    # differences of 1e-12, where the summation order of the two languages decides: there the objective must still agree.
    o40 = c_oracle.default_options(max_iter=40)
    c40 = c_oracle.solve_workload(w, nthreads=2, options=o40)
    o40.as_rounds = 0
    clip40 = c_oracle.solve_workload(w, nthreads=2, options=o40)
    for i, p in enumerate(problems_from_workload(w)):
        r = O.solve_lm(p, O.SolverOptions(max_iter=40))
        assert c40["status"][i] == r.status and c40["iters"][i] == r.iters
        np.testing.assert_allclose(c40["Q"][i], r.Q, atol=1e-8)
        r0 = O.solve_lm(p, O.SolverOptions(max_iter=40, as_rounds=0))
        assert clip40["iters"][i] == r0.iters
        np.testing.assert_allclose(clip40["Q"][i], r0.Q, atol=1e-8)
        rf = O.solve_lm(p)
        assert abs(with_as["cost"][i] - rf.cost) <= 1e-6 * rf.cost
        if with_as["status"][i] == 0 and rf.status == 0:
            np.testing.assert_allclose(with_as["Q"][i], rf.Q, atol=1e-6)
    # |dq| <= tol_step under heavy damping is "rests on a kink" (STATUS_SLOW), not converged
    loose = c_oracle.solve_workload(w, nthreads=2, options=c_oracle.default_options(lambda_conv=1e30))
    assert np.all((loose["status"] == 0) | (loose["status"] == with_as["status"]))
    assert np.all(with_as["status"][loose["status"] == 4] == 4)
