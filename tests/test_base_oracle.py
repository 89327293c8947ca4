"""CPU: the base-placement oracle (oracle/base_oracle.py) against finite differences and SciPy."""
import numpy as np

import base_oracle as BO
from helpers import base_placement_case


def _problem(t, qc, RTs, grid, origin, res, b=0):
    return BO.BaseProblem(t, qc, RTs[b], 0.01, grid, origin, res)


def test_jacobians_match_finite_differences():
    t, qc, RTs, grid, origin, res = base_placement_case("fetch_small", B=1, n=3)
    p = _problem(t, qc, RTs, grid, origin, res)
    y = np.array([0.1, -0.2, 0.3])
    Qx = np.tile(qc[t.opt_qidx], (p.n, 1)) + 0.02
    _, Jq, Jy = BO.residuals(p, y, Qx)
    h = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (BO.residuals(p, y + e, Qx, False)[0] - BO.residuals(p, y - e, Qx, False)[0]) / (2 * h)
        assert np.abs(fd - Jy[..., k]).max() < 1e-8
    for k in range(t.nopt):
        e = np.zeros_like(Qx); e[:, k] = h
        fd = (BO.residuals(p, y, Qx + e, False)[0] - BO.residuals(p, y, Qx - e, False)[0]) / (2 * h)
        assert np.abs(fd - Jq[..., k]).max() < 1e-8


def test_lm_reaches_the_scipy_optimum():
    """Where the LM iteration converges within the reference's max_iter = 100 it sits at SciPy's optimum.  The arm is redundant
    (the joint vectors of two solvers need not agree); the cost is unique and the base pose nearly so (flat valley: the effort
    weight is 0.01).  Problems that are still crawling along the valley after 100 iterations return their best iterate with
    STATUS_MAX_ITER -- the reference's IPOPT call stops at the same iteration cap (gto/base_planner.py:92)."""
    nconv = 0
    for name in ("panda_small", "fetch_small"):
        t, qc, RTs, grid, origin, res = base_placement_case(name, B=4, n=4, spread=0.5)
        for b in range(4):
            p = _problem(t, qc, RTs, grid, origin, res, b)
            r = BO.solve_base(p)
            assert r.status in (BO.STATUS_CONVERGED, BO.STATUS_MAX_ITER)
            assert abs(r.y[2]) <= np.pi
            assert np.all(r.Q[:, t.opt_qidx] >= t.lo - 1e-12) and np.all(r.Q[:, t.opt_qidx] <= t.hi + 1e-12)
            assert np.all(r.Q[:, t.par_qidx] == qc[t.par_qidx])
            assert r.cost <= BO.cost_of(p, np.zeros(3), np.tile(qc[t.opt_qidx], (p.n, 1)))
            if r.status == BO.STATUS_CONVERGED:
                nconv += 1
                ys, _, fs = BO.solve_scipy(p)
                assert abs(r.cost - fs) <= 1e-5 * max(fs, 1e-3), (r.cost, fs)
                assert np.abs(r.y - ys).max() < 2e-3, (r.y, ys)
    assert nconv >= 3


def test_collision_count_uses_reference_indexing():
    """floor((xy - origin)/res), clipped per axis, offset = iy + ny*ix (gto/gto_models.py:261-271)."""
    t, qc, RTs, grid, origin, res = base_placement_case("panda_small", B=1, n=2)
    p = _problem(t, qc, RTs, grid, origin, res)
    full = np.ones_like(grid)
    p.occupancy = full
    assert BO.collision_cost(p, np.zeros(3)) == t.npoints  # every point lands in some (clipped) cell
    p.occupancy = np.zeros_like(grid)
    assert BO.collision_cost(p, np.array([0.3, 0.1, 0.5])) == 0.0
