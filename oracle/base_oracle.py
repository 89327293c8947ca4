"""CPU oracle for mobile-base placement -- TEST INFRASTRUCTURE, not a product path.

Restates the reference's ``BasePlanner`` (``gto/base_planner.py:35-168``) in float64 NumPy.
Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of the benchmark tools may import this file.

Problem (one ``plan_goalset(qc, RTs)`` call, ``n`` goals):

    unknowns   y = (x, y, theta)                       TaskModel(dim=3), base_planner.py:23,44-51
               q_i, i < n                              one arm configuration per goal (builder T = goal_size, :38)
    cost       w_e |y|^2                               :57
             + sum_i sum_k | F(q_i) x_k - (T_b(y) RT_i G) x_k |^2          :70-86
               T_b(y) = [Rz(theta) | (x, y, 0)]  ("old base in new base", :49-52), F = frame of link_gripper,
               G = link_gripper in link_ee (constant), x_k = the gripper point set
    bounds     -pi <= theta <= pi (:54), lo <= q_i <= hi (:89); parameter joints stay at qc (:101-118)
    seed       q_i = qc, y = 0 (:101-102)

Afterwards (:131-165): per-goal position / rotation error and the occupancy-grid collision count of the robot (all surface
points at ``qc``) seen from the new base, ``sum grid[offset(RT_base^-1 W)]`` with the 2-D indexing of
``GTORobotModel.points_to_offsets_occupancy_numpy`` (``gto/gto_models.py:261-271``).

The reference hands this NLP to IPOPT; as for the trajectory problem (``gto_oracle.solve_lm``) the oracle and the CUDA kernel
solve it with the same projected Levenberg-Marquardt iteration, and ``tests/test_base_oracle.py`` pins the optimum against
SciPy's trust-region least squares on the literal per-point residuals.  Parity unpinned against the reference itself
(casadi/IPOPT absent, the reference stores no base-placement outputs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

import gto_oracle as O

STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_NAN, STATUS_STALLED = 0, 1, 2, 3


def base_tf(y) -> np.ndarray:
    """``rt2tr(rotz(theta), (x, y, 0))`` (base_planner.py:49-52)."""
    c, s = np.cos(y[2]), np.sin(y[2])
    T = np.eye(4)
    T[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    T[0, 3], T[1, 3] = y[0], y[1]
    return T


@dataclass
class BaseProblem:
    table: object  # RobotTable with link_ee / link_gripper set
    qc: np.ndarray  # [ndof]
    RTs: np.ndarray  # [n,4,4] goals in the current base frame
    w_effort: float = 0.01
    occupancy: Optional[np.ndarray] = None  # [nx,ny]
    occ_origin: Optional[np.ndarray] = None  # [2]
    occ_resolution: float = 0.05

    @property
    def n(self) -> int:
        return int(self.RTs.shape[0])

    def goal_frames(self) -> np.ndarray:
        """``RT_i . G`` as [n,4,4]."""
        G = O.hom(self.table.G)
        return np.stack([RT @ G for RT in self.RTs])

    def gripper_points(self) -> np.ndarray:
        tb = self.table
        return np.asarray(tb.points[tb.grip_pt_start : tb.grip_pt_start + tb.grip_pt_count], dtype=np.float64)


def full_q(p: BaseProblem, Qx: np.ndarray) -> np.ndarray:
    Q = np.tile(np.asarray(p.qc, dtype=np.float64).reshape(1, -1), (p.n, 1))
    Q[:, p.table.opt_qidx] = Qx
    return Q


def residuals(p: BaseProblem, y: np.ndarray, Qx: np.ndarray, need_jac: bool = True):
    """Literal per-point residuals ``F(q_i) x_k - T_b(y) RT_i G x_k`` [n,Pg,3] and their Jacobians w.r.t. q_i [n,Pg,3,nopt]
    and y [n,Pg,3,3] (the geometric Jacobian of optas/models.py:1203-1268 for the arm part)."""
    tb = p.table
    X = p.gripper_points()
    A = p.goal_frames()
    Tb = base_tf(y)
    c, s = np.cos(y[2]), np.sin(y[2])
    dRz = np.array([[-s, -c, 0.0], [c, -s, 0.0], [0.0, 0.0, 0.0]])
    Q = full_q(p, Qx)
    n, Pg, nopt = p.n, X.shape[0], tb.nopt
    r = np.zeros((n, Pg, 3))
    Jq = np.zeros((n, Pg, 3, nopt)) if need_jac else None
    Jy = np.zeros((n, Pg, 3, 3)) if need_jac else None
    for i in range(n):
        Tm = O.fk_movable(tb, Q[i])
        F = O.gripper_frame(tb, Q[i], Tm)
        W = X @ F[:3, :3].T + F[:3, 3]
        M = Tb @ A[i]
        r[i] = W - (X @ M[:3, :3].T + M[:3, 3])
        if need_jac:
            om, mm = O.joint_twists(tb, Tm)
            for k in range(nopt):
                if (tb.grip_optmask >> k) & 1:
                    Jq[i, :, :, k] = np.cross(om[k][None, :], W) + mm[k]
            Ai = X @ A[i][:3, :3].T + A[i][:3, 3]  # goal points before the base transform
            Jy[i, :, 0, 0] = -1.0
            Jy[i, :, 1, 1] = -1.0
            Jy[i, :, :, 2] = -(Ai @ dRz.T)
    return r, Jq, Jy


def cost_of(p: BaseProblem, y, Qx) -> float:
    r, _, _ = residuals(p, y, Qx, need_jac=False)
    return float(p.w_effort * np.dot(y, y) + np.sum(r * r))


def _system(p: BaseProblem, y, Qx):
    """Gauss-Newton arrow system: per-goal blocks H [n,nopt,nopt], couplings C [n,nopt,3], base block S [3,3],
    half gradients gq [n,nopt], gy [3], cost."""
    r, Jq, Jy = residuals(p, y, Qx)
    H = np.einsum("ipak,ipal->ikl", Jq, Jq)
    C = np.einsum("ipak,ipal->ikl", Jq, Jy)
    S = np.einsum("ipak,ipal->kl", Jy, Jy) + p.w_effort * np.eye(3)
    gq = np.einsum("ipak,ipa->ik", Jq, r)
    gy = np.einsum("ipak,ipa->k", Jy, r) + p.w_effort * np.asarray(y)
    F = float(p.w_effort * np.dot(y, y) + np.sum(r * r))
    return H, C, S, gq, gy, F


Y_LO = np.array([-1e30, -1e30, -np.pi])
Y_HI = np.array([1e30, 1e30, np.pi])


def lm_step(p: BaseProblem, y, Qx, sysm, lam, opts: O.SolverOptions):
    """One damped projected Gauss-Newton step on the arrow system, solved through the 3x3 Schur complement of the base block."""
    tb = p.table
    H, C, S, gq, gy, _ = sysm
    n, nopt = gq.shape
    fq = ((Qx <= tb.lo + opts.bound_eps) & (gq > 0)) | ((Qx >= tb.hi - opts.bound_eps) & (gq < 0))
    fy = ((y <= Y_LO + opts.bound_eps) & (gy > 0)) | ((y >= Y_HI - opts.bound_eps) & (gy < 0))
    for k in range(nopt):  # joints that do not move the gripper have a zero row: treat as fixed
        if not (tb.grip_optmask >> k) & 1:
            fq[:, k] = True
    pgq, pgy = np.where(fq, 0.0, gq), np.where(fy, 0.0, gy)
    pgnorm = 2.0 * max(float(np.max(np.abs(pgq))), float(np.max(np.abs(pgy))))
    if pgnorm <= opts.tol_grad:
        return None, None, 0.0, pgnorm, 0.0
    Sd = S + lam * np.diag(np.diag(S))
    Sd[fy, :] = 0.0
    Sd[:, fy] = 0.0
    Sd[fy, fy] = 1.0
    rhs_y = -pgy.copy()
    Z = np.zeros((n, nopt, 4))
    Cm = C.copy()
    for i in range(n):
        Hd = H[i] + lam * np.diag(np.diag(H[i]))
        f = fq[i]
        Hd[f, :] = 0.0
        Hd[:, f] = 0.0
        Hd[f, f] = 1.0
        Cm[i][f, :] = 0.0
        Cm[i][:, fy] = 0.0
        Z[i] = np.linalg.solve(Hd, np.concatenate([Cm[i], -pgq[i][:, None]], axis=1))
        Sd -= Cm[i].T @ Z[i][:, :3]
        rhs_y -= Cm[i].T @ Z[i][:, 3]
    dy = np.linalg.solve(Sd, rhs_y)
    dq = Z[:, :, 3] - np.einsum("ikl,l->ik", Z[:, :, :3], dy)
    yn = np.clip(y + dy, Y_LO, Y_HI)
    Qn = np.clip(Qx + dq, tb.lo, tb.hi)
    dy, dq = yn - y, Qn - Qx
    # predicted reduction with the undamped model
    Ad_q = np.einsum("ikl,il->ik", H, dq) + np.einsum("ikl,l->ik", C, dy)
    Ad_y = S @ dy + np.einsum("ikl,ik->l", C, dq)
    pred = -(float(np.sum(gq * dq)) + float(np.dot(gy, dy)) + 0.5 * (float(np.sum(dq * Ad_q)) + float(np.dot(dy, Ad_y))))
    step = max(float(np.max(np.abs(dq))), float(np.max(np.abs(dy))))
    return yn, Qn, pred, pgnorm, step


@dataclass
class BaseResult:
    Q: np.ndarray  # [n,ndof]
    y: np.ndarray  # [3]
    cost: float
    iters: int
    status: int
    err_pos: np.ndarray
    err_rot: np.ndarray
    collision: float


def pose_errors(p: BaseProblem, y, Q):
    """base_planner.py:131-148."""
    tb = p.table
    A = p.goal_frames()
    Tb = base_tf(y)
    ep, er = np.zeros(p.n), np.zeros(p.n)
    for i in range(p.n):
        RT = Tb @ A[i]
        tf = O.gripper_frame(tb, Q[i])
        ep[i] = np.linalg.norm(RT[:3, 3] - tf[:3, 3])
        # angle between the two rotations; equals arccos(2 <q1,q2>^2 - 1) of the reference's quaternion form
        cosang = 0.5 * (np.trace(RT[:3, :3].T @ tf[:3, :3]) - 1.0)
        er[i] = np.degrees(np.arccos(np.clip(cosang, -1.0, 1.0)))
    return ep, er


def collision_cost(p: BaseProblem, y) -> float:
    """base_planner.py:150-165 with gto_models.py:261-271."""
    if p.occupancy is None:
        return 0.0
    W = O.world_points(p.table, p.qc)
    Tinv = np.linalg.inv(base_tf(y))
    Wn = W @ Tinv[:3, :3].T + Tinv[:3, 3]
    nx, ny = p.occupancy.shape
    idx = np.floor((Wn[:, :2] - np.asarray(p.occ_origin).reshape(1, 2)) / p.occ_resolution)
    ix = np.clip(idx[:, 0], 0, nx - 1).astype(np.int64)
    iy = np.clip(idx[:, 1], 0, ny - 1).astype(np.int64)
    return float(np.sum(p.occupancy.reshape(-1)[iy + ny * ix]))


def solve_base(p: BaseProblem, opts: Optional[O.SolverOptions] = None) -> BaseResult:
    """Projected LM with the damping policy of ``gto_oracle.solve_lm`` (no fp32 noise floor: everything is float64)."""
    opts = opts or O.SolverOptions()
    tb = p.table
    y = np.zeros(3)
    Qx = np.tile(np.asarray(p.qc, dtype=np.float64)[tb.opt_qidx].reshape(1, -1), (p.n, 1))
    sysm = _system(p, y, Qx)
    F = sysm[5]
    lam, nu = opts.lambda0, 2.0
    status, it = STATUS_MAX_ITER, 0
    while it < opts.max_iter:
        yn, Qn, pred, pgnorm, step = lm_step(p, y, Qx, sysm, lam, opts)
        if yn is None:
            status = STATUS_CONVERGED
            break
        it += 1
        sys_t = _system(p, yn, Qn)
        Ft = sys_t[5]
        if not np.isfinite(Ft):
            status = STATUS_NAN
            break
        ared = 0.5 * (F - Ft)
        if pred > 0 and ared >= opts.eta * pred:
            rho = ared / pred
            y, Qx, sysm, F = yn, Qn, sys_t, Ft
            lam = max(opts.lambda_min, lam * max(1.0 / 3.0, 1.0 - (2.0 * min(rho, 1.0) - 1.0) ** 3))
            nu = 2.0
            if step <= opts.tol_step:
                status = STATUS_CONVERGED
                break
        else:
            if pred <= 0 and step <= opts.tol_step:
                status = STATUS_CONVERGED
                break
            lam = min(opts.lambda_max, lam * nu)
            nu *= 2.0
            if lam >= opts.lambda_max:
                status = STATUS_STALLED
                break
    Q = full_q(p, Qx)
    ep, er = pose_errors(p, y, Q)
    return BaseResult(Q, y, F, it, status, ep, er, collision_cost(p, y))


def solve_scipy(p: BaseProblem):
    """Independent check: SciPy trust-region-reflective least squares on the literal residual vector."""
    from scipy.optimize import least_squares

    tb = p.table
    n, nopt = p.n, tb.nopt

    def unpack(x):
        return x[:3], x[3:].reshape(n, nopt)

    def fun(x):
        y, Qx = unpack(x)
        r, _, _ = residuals(p, y, Qx, need_jac=False)
        return np.concatenate([np.sqrt(p.w_effort) * y, r.reshape(-1)])

    def jac(x):
        y, Qx = unpack(x)
        r, Jq, Jy = residuals(p, y, Qx)
        Pg = r.shape[1]
        J = np.zeros((3 + n * Pg * 3, 3 + n * nopt))
        J[:3, :3] = np.sqrt(p.w_effort) * np.eye(3)
        for i in range(n):
            rows = slice(3 + i * Pg * 3, 3 + (i + 1) * Pg * 3)
            J[rows, :3] = Jy[i].reshape(Pg * 3, 3)
            J[rows, 3 + i * nopt : 3 + (i + 1) * nopt] = Jq[i].reshape(Pg * 3, nopt)
        return J

    x0 = np.concatenate([np.zeros(3), np.tile(np.asarray(p.qc, dtype=np.float64)[tb.opt_qidx], n)])
    lo = np.concatenate([Y_LO, np.tile(tb.lo, n)])
    hi = np.concatenate([Y_HI, np.tile(tb.hi, n)])
    lo[:2], hi[:2] = -np.inf, np.inf
    x0 = np.clip(x0, lo + 1e-9, hi - 1e-9)
    sol = least_squares(fun, x0, jac=jac, bounds=(lo, hi), xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=2000)
    y, Qx = unpack(sol.x)
    return y, Qx, float(np.sum(fun(sol.x) ** 2))
