"""Self-consistency of the oracle: analytic Jacobian vs finite differences, projected LM vs SciPy TRF, bounds,
equivalence of the reduced problem with the reference's full-space constraints.  CPU only."""
import numpy as np
import pytest

import gto_oracle as O
from helpers import problems_from_workload, small_workload


@pytest.fixture(scope="module")
def wl():
    return small_workload("C2", "panda_small", B=3, n_field=48)


def test_jacobian_rows_match_finite_differences(wl):
    p = problems_from_workload(wl, [1])[0]
    Q = O.initial_trajectory(p)
    # move the last knots to the goal configuration so that the fingers are inside the obstacle shell
    Q[-6:, p.table.opt_qidx] = wl.q_star[1][p.table.opt_qidx]
    lin = O.linearize(p, Q)
    assert np.abs(lin.J_obs).max() > 0
    h = 1e-6
    for t in (p.knot_standoff, p.T - 1):
        for k in range(p.table.nopt):
            Qp, Qm = Q.copy(), Q.copy()
            Qp[t, p.table.opt_qidx[k]] += h
            Qm[t, p.table.opt_qidx[k]] -= h
            lp, lm = O.linearize(p, Qp, need_jac=False), O.linearize(p, Qm, need_jac=False)
            fd = (lp.r_obs[t] - lm.r_obs[t]) / (2 * h)
            smooth = np.abs(fd - lin.J_obs[t, :, k]) < 1e-4  # points that cross a voxel face inside +-h are excluded
            assert smooth.mean() > 0.99
            if t == p.T - 1:
                np.testing.assert_allclose((lp.r_goal - lm.r_goal) / (2 * h), lin.J_goal[:, :, k], atol=1e-7)
            else:
                np.testing.assert_allclose((lp.r_stand - lm.r_stand) / (2 * h), lin.J_stand[:, :, k], atol=1e-7)


def test_block_tridiagonal_solver_matches_dense():
    rng = np.random.default_rng(0)
    m, n, off = 9, 5, -0.7
    D = np.zeros((m, n, n))
    for i in range(m):
        A = rng.normal(size=(n, n))
        D[i] = A @ A.T + 3 * np.eye(n)
    rhs = rng.normal(size=(m, n))
    fixed = rng.random((m, n)) < 0.2
    for i in range(m):
        f = fixed[i]
        D[i][f, :] = 0
        D[i][:, f] = 0
        D[i][f, f] = 1
    rhs[fixed] = 0
    x = O._solve_masked(D, off, rhs, fixed)
    full = np.zeros((m * n, m * n))
    for i in range(m):
        full[i * n:(i + 1) * n, i * n:(i + 1) * n] = D[i]
        if i + 1 < m:
            C = off * np.diag((~fixed[i] & ~fixed[i + 1]).astype(float))
            full[(i + 1) * n:(i + 2) * n, i * n:(i + 1) * n] = C
            full[i * n:(i + 1) * n, (i + 1) * n:(i + 2) * n] = C
    np.testing.assert_allclose(x.reshape(-1), np.linalg.solve(full, rhs.reshape(-1)), atol=1e-10)


def test_lm_agrees_with_scipy_trf_zero_field(wl):
    p = problems_from_workload(wl, [0])[0]
    p.field_all = p.field_obs = None
    r = O.solve_lm(p)
    Qs, cs, _ = O.solve_scipy(p)
    assert r.status == O.STATUS_CONVERGED
    assert np.abs(r.Q - Qs).max() < 1e-4
    assert r.cost == pytest.approx(cs, rel=1e-6)


def test_solution_respects_constraints_and_bounds(wl):
    p = problems_from_workload(wl, [2])[0]
    p.T, p.dt, p.standoff_offset = 12, 10.0 / 11, -4  # short horizon keeps the SciPy cross-check fast
    p.q_seed = p.q_seed[:12]
    p.field_all = p.field_obs = None
    t = p.table
    # shrink every joint range to a box around qc: the goal becomes unreachable and bounds must become active
    import copy
    t2 = copy.deepcopy(t)
    t2.lo = np.maximum(t.lo, p.qc[t.opt_qidx] - 0.25)
    t2.hi = np.minimum(t.hi, p.qc[t.opt_qidx] + 0.25)
    p.table = t2
    r = O.solve_lm(p)
    Qx = r.Q[:, t.opt_qidx]
    assert np.all(Qx >= t2.lo - 1e-12) and np.all(Qx <= t2.hi + 1e-12)
    assert np.any(np.isclose(Qx, t2.lo)) or np.any(np.isclose(Qx, t2.hi))  # some bound is active
    np.testing.assert_array_equal(Qx[0], p.qc[t.opt_qidx])
    np.testing.assert_array_equal(Qx[1], p.qc[t.opt_qidx])
    np.testing.assert_allclose(r.dQ[:, t.opt_qidx], np.diff(Qx, axis=0) / p.dt)
    Qs, cs, _ = O.solve_scipy(p)
    assert r.cost == pytest.approx(cs, rel=1e-3)


def test_reduced_problem_equals_full_space_problem():
    """A11: eliminating dQ and the first two knots is exact.  Tiny instance solved in the reference's full space
    x = [vec(Q_x); vec(dQ_x)] with the reference's constraint rows (SLSQP) vs the reduced LM."""
    from scipy.optimize import minimize

    w = small_workload("C2", "panda_small", B=1, n_field=32)
    p = problems_from_workload(w, [0])[0]
    p.T, p.dt, p.standoff_offset = 6, 10.0 / 5, -2
    p.q_seed = p.q_seed[:6]
    p.field_all = p.field_obs = None
    t = p.table
    oi, n, T = t.opt_qidx, t.nopt, p.T
    r = O.solve_lm(p, O.SolverOptions(tol_grad=1e-9, tol_step=1e-9))
    Q0 = O.initial_trajectory(p)

    def split(x):
        return x[: n * T].reshape(T, n), x[n * T:].reshape(T - 1, n)

    def f(x):
        Qx, dQ = split(x)
        Q = Q0.copy()
        Q[:, oi] = Qx
        lin = O.linearize(p, Q, need_jac=False)
        return float(np.sum(lin.cost_pts)) + p.w_vel * float(np.sum(dQ * dQ))

    cons = [
        {"type": "eq", "fun": lambda x: split(x)[0][0] - p.qc[oi]},
        {"type": "eq", "fun": lambda x: split(x)[1][0]},
        {"type": "eq", "fun": lambda x: (split(x)[0][:-1] + p.dt * split(x)[1] - split(x)[0][1:]).reshape(-1)},
    ]
    bounds = [(lo, hi) for _ in range(T) for lo, hi in zip(t.lo, t.hi)] + [(None, None)] * (n * (T - 1))
    x0 = np.concatenate([r.Q[:, oi].reshape(-1) + 1e-3, r.dQ[:, oi].reshape(-1)])
    sol = minimize(f, x0, method="SLSQP", constraints=cons, bounds=bounds, options={"ftol": 1e-14, "maxiter": 300})
    Qx, dQ = split(sol.x)
    assert sol.fun == pytest.approx(r.cost, rel=1e-5)
    assert np.abs(Qx - r.Q[:, oi]).max() < 2e-3


def test_plan_goalset_semantics_min_over_goals(wl):
    """Q11: the reference's mmin over per-goal costs equals the best of the per-goal problems."""
    ps = problems_from_workload(wl)
    costs = [O.solve_lm(p).cost for p in ps]
    assert np.isfinite(costs).all() and min(costs) > 0
