"""CPU oracle for the batched grasp-trajectory hot path (float64 NumPy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``grasptrajopt_b200/`` may import this
module; it is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker / timed
CPU baseline, never as the shipped path.

It restates, function by function, the arithmetic of the reference's path
(IRVLUTD/GraspTrajOpt @ 4703ba2; all citations relative to the reference root):

=====  ==========================================  ======================================
row    reference                                   here
=====  ==========================================  ======================================
A1     optas/models.py:826-868 (chain FK)          ``fk_movable`` / ``link_frames``
       optas/spatialmath.py:91-100,187-225
A2     gto/gto_models.py:83-101 (visual frames)    ``link_frames``
A4     gto/gto_planner.py:111-128                  ``world_points``
A5     gto/gto_models.py:174-201 (nearest node)    ``Field.nearest`` / ``Field.offsets``
A6     gto/sdf_callback.py:38-49,90-114,165-183    ``Field.nearest``, ``Field.central_grad``,
                                                   ``Field.central_hess``
A6'    SURVEY.md Appendix A (trilinear, build)     ``Field.trilinear``
A7     mesh_to_sdf/depth_point_cloud.py:65-91      ``sdf_cost_transform``
A8     gto/gto_planner.py:86-105                   ``linearize`` (goal / stand-off rows)
A9     gto/gto_planner.py:131                      ``linearize`` (obstacle rows)
A10    gto/gto_planner.py:134-135                  ``velocity_terms``
A11    gto/gto_planner.py:59-72,138 +              eliminated analytically (``free knots``),
       optas/builder.py:420-524                    bounds kept (``solve_lm`` projection)
A13    optas/solver.py:335-400 (IPOPT)             ``solve_lm`` / ``lm_step`` (projected bundle
                                                   Levenberg-Marquardt: cutting planes from
                                                   rejected trial points for the kinks of the
                                                   trilinear field; same algorithm as the CUDA
                                                   solver and gto_oracle.c) and ``solve_scipy``
                                                   (SciPy TRF, independent cross-check)
A14    gto/utils.py:63-82, gto_planner.py:193-219  ``interpolate_seed``, ``plan_cost_nearest``
A15    optas/solver.py:126-159                     ``unpack_solution``
=====  ==========================================  ======================================

PARITY PINNING.  The reference itself cannot be imported here (``casadi``/IPOPT,
``urdf_parser_py``, ``trimesh`` are absent and there is no network).  Rows A1, A2, A5, A6,
A7, A14 are pinned against the reference's *own Python code* executed under numeric
stand-ins for its missing third-party imports (``oracle/refshim``; generator
``oracle/make_golden.py``; vectors in ``tests/golden/ref_*.npz``).  The NLP solve itself
(A13: CasADi AD + IPOPT) has no runnable reference and no stored input/output pair:
for that row **parity is unpinned**; it is anchored on (i) agreement of two independent
solvers here (projected LM vs SciPy TRF) and (ii) properties of the plans the reference
stored in ``examples/results_iros2024`` (``tests/golden/ref_stored_plans.npz``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

JOINT_REVOLUTE = 1
JOINT_PRISMATIC = 2


# --------------------------------------------------------------------------------------
# spatial primitives (optas/spatialmath.py:91-100, 187-225)
# --------------------------------------------------------------------------------------
def rodrigues(theta: float, axis: np.ndarray) -> np.ndarray:
    a = axis / np.linalg.norm(axis)
    K = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
    return np.eye(3) + np.sin(theta) * K + (1.0 - np.cos(theta)) * (K @ K)


def rpy_to_R(rpy) -> np.ndarray:
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rz = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    Ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    Rx = np.array([[1.0, 0.0, 0.0], [0.0, cr, -sr], [0.0, sr, cr]])
    return Rz @ Ry @ Rx


def hom(T34: np.ndarray) -> np.ndarray:
    T = np.eye(4)
    T[:3] = T34
    return T


# --------------------------------------------------------------------------------------
# A1/A2: forward kinematics on the flattened robot table
# --------------------------------------------------------------------------------------
def fk_movable(table, q: np.ndarray) -> np.ndarray:
    """World (robot-base) frame of every movable joint, *after* its motion: [nmov,4,4].

    ``T_j = T_parent(j) . origin_j . motion_j(q_j)`` with Rodrigues rotation for revolute /
    continuous joints and a translation ``q_j * axis`` for prismatic ones
    (optas/models.py:845-868)."""
    q = np.asarray(q, dtype=np.float64).reshape(-1)
    T = np.zeros((table.nmov, 4, 4))
    for j in range(table.nmov):
        P = np.eye(4) if table.mov_parent[j] < 0 else T[table.mov_parent[j]]
        M = np.eye(4)
        qj = q[table.mov_qidx[j]]
        if table.mov_type[j] == JOINT_REVOLUTE:
            M[:3, :3] = rodrigues(qj, table.mov_axis[j])
        else:
            M[:3, 3] = qj * table.mov_axis[j]
        T[j] = P @ hom(table.mov_origin[j]) @ M
    return T


def link_frames(table, q: np.ndarray, Tm: Optional[np.ndarray] = None) -> np.ndarray:
    """Visual frame of every collision link in the robot base frame: [nlinks,4,4]
    (gto/gto_models.py:92-100: ``link_tf(q) @ rt2tr(rpy2r(rpy_vis), xyz_vis)``)."""
    if Tm is None:
        Tm = fk_movable(table, q)
    out = np.zeros((table.nlinks, 4, 4))
    for l in range(table.nlinks):
        P = np.eye(4) if table.link_mov[l] < 0 else Tm[table.link_mov[l]]
        out[l] = P @ hom(table.link_tf[l])
    return out


def gripper_frame(table, q: np.ndarray, Tm: Optional[np.ndarray] = None) -> np.ndarray:
    """Plain link frame of ``link_gripper`` (gto/gto_planner.py:79-82)."""
    if Tm is None:
        Tm = fk_movable(table, q)
    P = np.eye(4) if table.grip_mov < 0 else Tm[table.grip_mov]
    return P @ hom(table.grip_tf)


def joint_twists(table, Tm: np.ndarray):
    """Per optimised joint k: (omega_k, m_k) such that the velocity of a base-frame point W
    under unit joint rate is ``omega_k x W + m_k``.

    Restates the geometric Jacobian of optas/models.py:1203-1268: revolute
    ``z x (W - o)`` = ``z x W + o x z``; prismatic ``z``."""
    om = np.zeros((table.nopt, 3))
    mm = np.zeros((table.nopt, 3))
    for j in range(table.nmov):
        k = table.mov_opt[j]
        if k < 0:
            continue
        z = Tm[j, :3, :3] @ table.mov_axis[j]
        o = Tm[j, :3, 3]
        if table.mov_type[j] == JOINT_REVOLUTE:
            om[k] = z
            mm[k] = np.cross(o, z)
        else:
            mm[k] = z
    return om, mm


def world_points(table, q, base_position=(0.0, 0.0, 0.0)):
    """A4: ``W = R_link(q) x + t_link(q) + base_position`` for every surface point
    (gto/gto_planner.py:113-116; numeric twin gto/gto_models.py:104-121)."""
    F = link_frames(table, q)
    W = np.zeros((table.npoints, 3))
    for l in range(table.nlinks):
        s, c = table.link_pt_start[l], table.link_pt_count[l]
        W[s : s + c] = table.points[s : s + c] @ F[l, :3, :3].T + F[l, :3, 3]
    return W + np.asarray(base_position, dtype=np.float64).reshape(1, 3)


# --------------------------------------------------------------------------------------
# A5/A6/A7: voxel cost field
# --------------------------------------------------------------------------------------
def sdf_cost_transform(distances: np.ndarray, epsilon: float = 0.02, w_inside: float = 1.0) -> np.ndarray:
    """A7 (mesh_to_sdf/depth_point_cloud.py:84-89): inside (d<0): ``w*(-d + eps/2)``;
    ``0<d<eps``: ``(d-eps)^2/(2 eps)``; otherwise 0.  float32 like the reference (:68)."""
    d = np.asarray(distances, dtype=np.float32).reshape(-1)
    inside = d < 0
    cost = np.zeros_like(d)
    cost[inside] = w_inside * (-d[inside] + epsilon / 2)
    shell = (d > 0) & (d < epsilon)
    cost[shell] = np.square(d[shell] - epsilon) / (2 * epsilon)
    return cost


@dataclass
class Field:
    """Node-centred voxel field; nodes at ``origin + k*pitch``; C-order flat index
    ``iz + Nz*(iy + Ny*ix)`` (gto/gto_models.py:155-187)."""

    cost: np.ndarray  # [Nx,Ny,Nz]
    origin: np.ndarray  # [3]
    pitch: float

    def __post_init__(self):
        self.cost = np.asarray(self.cost)
        assert self.cost.ndim == 3
        self.origin = np.asarray(self.origin, dtype=np.float64).reshape(3)
        self.pitch = float(self.pitch)

    @property
    def shape(self):
        return self.cost.shape

    # A5 symbolic form: floor then clamp (gto_models.py:174-187; sdf_callback.py:38-41)
    def indices(self, W: np.ndarray) -> np.ndarray:
        idx = np.floor((np.asarray(W, dtype=np.float64).reshape(-1, 3) - self.origin) / self.pitch).astype(np.int64)
        return np.clip(idx, 0, np.array(self.shape) - 1)

    def offsets(self, W: np.ndarray) -> np.ndarray:
        i = self.indices(W)
        return i[:, 2] + self.shape[2] * (i[:, 1] + self.shape[1] * i[:, 0])

    # A5 numeric twin used for seed ranking: clip THEN truncate (gto_models.py:190-201)
    def offsets_numpy_twin(self, W: np.ndarray) -> np.ndarray:
        u = (np.asarray(W, dtype=np.float64).reshape(-1, 3) - self.origin) / self.pitch
        i = np.stack([np.clip(u[:, a], 0, self.shape[a] - 1).astype(np.int32) for a in range(3)], axis=1).astype(np.int64)
        off = i[:, 2] + self.shape[2] * (i[:, 1] + self.shape[1] * i[:, 0])
        return np.clip(off, 0, self.cost.size - 1)

    def nearest(self, W: np.ndarray) -> np.ndarray:
        return self.cost.reshape(-1)[self.offsets(W)].astype(np.float64)

    # A6: central differences over +-1 clamped voxel (sdf_callback.py:90-114)
    def central_grad(self, W: np.ndarray) -> np.ndarray:
        i = self.indices(W)
        hi = np.array(self.shape) - 1
        g = np.zeros((i.shape[0], 3))
        for a in range(3):
            ip, im = i.copy(), i.copy()
            ip[:, a] = np.minimum(i[:, a] + 1, hi[a])
            im[:, a] = np.maximum(i[:, a] - 1, 0)
            g[:, a] = (self.cost[ip[:, 0], ip[:, 1], ip[:, 2]].astype(np.float64) - self.cost[im[:, 0], im[:, 1], im[:, 2]]) / (2 * self.pitch)
        return g

    # A6: 4-point mixed differences (sdf_callback.py:165-183)
    def central_hess(self, W: np.ndarray) -> np.ndarray:
        i = self.indices(W)
        hi = np.array(self.shape) - 1
        n = i.shape[0]
        H = np.zeros((n, 3, 3))

        def val(j):
            j = np.clip(j, 0, hi)
            return self.cost[j[:, 0], j[:, 1], j[:, 2]].astype(np.float64)

        E = np.eye(3, dtype=np.int64)
        for a in range(3):
            for b in range(a, 3):
                d = (val(i + E[a] + E[b]) - val(i + E[a] - E[b]) - val(i - E[a] + E[b]) + val(i - E[a] - E[b])) / (4 * self.pitch**2)
                H[:, a, b] = d
                H[:, b, a] = d
        return H

    # A6': trilinear value + analytic gradient -- what the kernels compute
    def trilinear(self, W: np.ndarray):
        W = np.asarray(W, dtype=np.float64).reshape(-1, 3)
        N = np.array(self.shape)
        u = (W - self.origin) / self.pitch
        i = np.clip(np.floor(u), 0, N - 2).astype(np.int64)
        f = u - i
        inb = (f >= 0.0) & (f <= 1.0)  # gradient is zero along an axis where f is clamped
        f = np.clip(f, 0.0, 1.0)
        C = self.cost.astype(np.float64)
        ix, iy, iz = i[:, 0], i[:, 1], i[:, 2]
        c = {}
        for a in (0, 1):
            for b in (0, 1):
                for d in (0, 1):
                    c[a, b, d] = C[ix + a, iy + b, iz + d]
        fx, fy, fz = f[:, 0], f[:, 1], f[:, 2]
        # interpolate along z, then y, then x (the kernels use the same association order)
        cz = {(a, b): c[a, b, 0] + fz * (c[a, b, 1] - c[a, b, 0]) for a in (0, 1) for b in (0, 1)}
        dz = {(a, b): c[a, b, 1] - c[a, b, 0] for a in (0, 1) for b in (0, 1)}
        cy = {a: cz[a, 0] + fy * (cz[a, 1] - cz[a, 0]) for a in (0, 1)}
        val = cy[0] + fx * (cy[1] - cy[0])
        gx = cy[1] - cy[0]
        dy0 = cz[0, 1] - cz[0, 0]
        dy1 = cz[1, 1] - cz[1, 0]
        gy = dy0 + fx * (dy1 - dy0)
        dzy0 = dz[0, 0] + fy * (dz[0, 1] - dz[0, 0])
        dzy1 = dz[1, 0] + fy * (dz[1, 1] - dz[1, 0])
        gz = dzy0 + fx * (dzy1 - dzy0)
        grad = np.stack([gx, gy, gz], axis=1) / self.pitch
        grad = np.where(inb, grad, 0.0)
        return val, grad


def zero_field() -> Field:
    return Field(np.zeros((2, 2, 2), dtype=np.float32), np.zeros(3), 1.0)


# --------------------------------------------------------------------------------------
# problem definition
# --------------------------------------------------------------------------------------
@dataclass
class Problem:
    """One (seed x grasp) trajectory problem == one reference ``plan()`` call
    (gto/gto_planner.py:42-142)."""

    table: object
    T: int
    dt: float
    qc: np.ndarray  # [ndof]
    RT: np.ndarray  # [4,4] goal pose of link_ee in the robot base frame
    q_seed: np.ndarray  # [T,ndof]
    base_position: np.ndarray = field(default_factory=lambda: np.zeros(3))
    field_all: Optional[Field] = None  # knots < T+standoff_offset   (gto_planner.py:117-121)
    field_obs: Optional[Field] = None  # knots >= T+standoff_offset  (gto_planner.py:122-126)
    standoff_offset: int = -10
    standoff_distance: float = -0.1
    axis_standoff: str = "x"
    use_standoff: bool = True
    collision_avoidance: bool = True
    w_goal: float = 1.0
    w_obs: float = 10.0
    w_vel: float = 0.01
    sdf_mode: str = "trilinear"  # "trilinear" (build) | "nearest" (A5, zero gradient)
    obs_linear: bool = False  # obstacle term w_obs * sum c instead of w_obs * sum c^2 (the IK solver's term, gto/ik_solver.py:69)

    def __post_init__(self):
        self.qc = np.asarray(self.qc, dtype=np.float64).reshape(-1)
        self.RT = np.asarray(self.RT, dtype=np.float64).reshape(4, 4)
        self.q_seed = np.asarray(self.q_seed, dtype=np.float64).reshape(self.T, -1)
        self.base_position = np.asarray(self.base_position, dtype=np.float64).reshape(3)

    @property
    def knot_standoff(self) -> int:
        return self.T + self.standoff_offset

    def goal_targets(self):
        """Target positions of the gripper points at the last knot and at the stand-off knot:
        ``(RT.G) x_k`` and ``(RT.S.G) x_k`` (gto_planner.py:93-102)."""
        t = self.table
        x = t.points[t.grip_pt_start : t.grip_pt_start + t.grip_pt_count]
        G = hom(t.G)
        M = self.RT @ G
        S = np.eye(4)
        S["xyz".index(self.axis_standoff), 3] = self.standoff_distance
        Ms = self.RT @ S @ G
        return x @ M[:3, :3].T + M[:3, 3], x @ Ms[:3, :3].T + Ms[:3, 3]

    def field_for_knot(self, t: int) -> Optional[Field]:
        return self.field_all if t < self.knot_standoff else self.field_obs


def initial_trajectory(p: Problem) -> np.ndarray:
    """Seed projected on the constraints of A11: optimised rows of knots 0 and 1 equal ``qc``
    (``Q[:,0]=qc`` and ``dQ[:,0]=0``, gto_planner.py:59-72), optimised rows clipped to the
    position limits (:138); parameter-joint rows are kept from the seed (``<robot>/q/p``,
    gto_planner.py:175,235)."""
    t = p.table
    Q = p.q_seed.copy()
    Q[:, t.opt_qidx] = np.clip(Q[:, t.opt_qidx], t.lo, t.hi)
    Q[0, t.opt_qidx] = p.qc[t.opt_qidx]
    Q[1, t.opt_qidx] = p.qc[t.opt_qidx]
    return Q


# --------------------------------------------------------------------------------------
# A8/A9: residual rows and analytic Jacobian rows
# --------------------------------------------------------------------------------------
@dataclass
class Linearization:
    r_obs: np.ndarray  # [T,P]       sqrt(w_obs) * c(W)
    J_obs: np.ndarray  # [T,P,nopt]  d r_obs / d Q_t
    r_goal: np.ndarray  # [Pg,3]
    J_goal: np.ndarray  # [Pg,3,nopt] wrt Q_{T-1}
    r_stand: np.ndarray  # [Pg,3]
    J_stand: np.ndarray  # [Pg,3,nopt] wrt Q_{T+so}
    H: np.ndarray  # [T,nopt,nopt]  sum_rows j j^T per knot
    g: np.ndarray  # [T,nopt]       sum_rows j r   per knot
    cost_pts: np.ndarray  # [T]           sum_rows r^2   per knot


def linearize(p: Problem, Q: np.ndarray, need_jac: bool = True) -> Linearization:
    tb = p.table
    T, P, n = p.T, tb.npoints, tb.nopt
    Pg = tb.grip_pt_count
    sw = np.sqrt(p.w_obs)
    sg = np.sqrt(p.w_goal)
    r_obs = np.zeros((T, P))
    J_obs = np.zeros((T, P, n))
    r_goal = np.zeros((Pg, 3))
    J_goal = np.zeros((Pg, 3, n))
    r_stand = np.zeros((Pg, 3))
    J_stand = np.zeros((Pg, 3, n))
    H = np.zeros((T, n, n))
    g = np.zeros((T, n))
    cost = np.zeros(T)
    tgt_goal, tgt_stand = p.goal_targets()
    xg = tb.points[tb.grip_pt_start : tb.grip_pt_start + Pg]
    masks = [np.array([(int(m) >> k) & 1 for k in range(n)], dtype=np.float64) for m in tb.link_optmask]
    gmask = np.array([(int(tb.grip_optmask) >> k) & 1 for k in range(n)], dtype=np.float64)

    for t in range(T):
        q = Q[t]
        Tm = fk_movable(tb, q)
        om, mm = joint_twists(tb, Tm)
        if p.collision_avoidance:
            fld = p.field_for_knot(t)
            if fld is not None:
                F = link_frames(tb, q, Tm)
                for l in range(tb.nlinks):
                    s, c = tb.link_pt_start[l], tb.link_pt_count[l]
                    Wb = tb.points[s : s + c] @ F[l, :3, :3].T + F[l, :3, 3]  # base frame
                    Ww = Wb + p.base_position
                    if p.sdf_mode == "trilinear":
                        val, grad = fld.trilinear(Ww)
                    else:
                        val, grad = fld.nearest(Ww), np.zeros((c, 3))
                    r_obs[t, s : s + c] = sw * val
                    if need_jac:
                        # row_k = grad . (om_k x W + m_k) = om_k . (W x grad) + m_k . grad
                        nvec = np.cross(Wb, grad)
                        J_obs[t, s : s + c] = sw * (nvec @ om.T + grad @ mm.T) * masks[l]
        for which, knot, tgt, rr, JJ in (("goal", T - 1, tgt_goal, r_goal, J_goal), ("stand", p.knot_standoff, tgt_stand, r_stand, J_stand)):
            if knot != t or (which == "stand" and not p.use_standoff):
                continue
            Fg = gripper_frame(tb, q, Tm)
            Wg = xg @ Fg[:3, :3].T + Fg[:3, 3]
            rr[:] = sg * (Wg - tgt)
            if need_jac:
                # d W / d q_k = om_k x W + m_k   -> [Pg,3,nopt]
                vel = np.cross(om[None, :, :], Wg[:, None, :]) + mm[None, :, :]  # [Pg,nopt,3]
                JJ[:] = sg * np.transpose(vel, (0, 2, 1)) * gmask
        rows_r = [r_obs[t]]
        rows_J = [J_obs[t]]
        if t == T - 1:
            rows_r.append(r_goal.reshape(-1))
            rows_J.append(J_goal.reshape(-1, n))
        if t == p.knot_standoff and p.use_standoff:
            rows_r.append(r_stand.reshape(-1))
            rows_J.append(J_stand.reshape(-1, n))
        if p.obs_linear:
            # unsquared obstacle term w*sum(c) (gto/ik_solver.py:69): value w*c, half gradient (w/2) dc/dq, no Gauss-Newton
            # curvature (c is piecewise trilinear).  r_obs / J_obs still hold sqrt(w)*c and sqrt(w)*dc/dq.
            rows_r, rows_J = rows_r[1:], rows_J[1:]
        rr_ = np.concatenate(rows_r) if rows_r else np.zeros(0)
        JJ_ = np.concatenate(rows_J, axis=0) if rows_J else np.zeros((0, n))
        cost[t] = rr_ @ rr_
        if need_jac:
            H[t] = JJ_.T @ JJ_
            g[t] = JJ_.T @ rr_
        if p.obs_linear:
            cost[t] += sw * float(np.sum(r_obs[t]))
            if need_jac:
                g[t] += 0.5 * sw * J_obs[t].sum(axis=0)
    return Linearization(r_obs, J_obs, r_goal, J_goal, r_stand, J_stand, H, g, cost)


def pack_rows(p: Problem, lin: Linearization) -> np.ndarray:
    """Dense Jacobian rows in the C-ABI layout ``[rows][nopt+1]`` = ``[J | r]`` with rows
    ordered: obstacle ``[t][point]`` (if collision_avoidance), goal ``[axis][k]``, stand-off
    ``[axis][k]`` (include/gto_b200.h, ``gto_eval_batch``)."""
    n = p.table.nopt
    blocks = []
    if p.collision_avoidance:
        blocks.append(np.concatenate([lin.J_obs.reshape(-1, n), lin.r_obs.reshape(-1, 1)], axis=1))
    blocks.append(np.concatenate([np.transpose(lin.J_goal, (1, 0, 2)).reshape(-1, n), lin.r_goal.T.reshape(-1, 1)], axis=1))
    if p.use_standoff:
        blocks.append(np.concatenate([np.transpose(lin.J_stand, (1, 0, 2)).reshape(-1, n), lin.r_stand.T.reshape(-1, 1)], axis=1))
    return np.concatenate(blocks, axis=0)


# --------------------------------------------------------------------------------------
# A10: velocity regulariser; total cost
# --------------------------------------------------------------------------------------
def velocity_cost(p: Problem, Qx: np.ndarray) -> float:
    """``w_vel * sum dQ^2`` with ``dQ_t = (Q_{t+1}-Q_t)/dt`` over the optimised rows
    (gto_planner.py:134-135; parameter rows of dQ are zero parameters)."""
    d = np.diff(Qx, axis=0) / p.dt
    return float(p.w_vel * np.sum(d * d))


def total_cost(p: Problem, Q: np.ndarray, lin: Optional[Linearization] = None) -> float:
    """Reference objective ``f`` = goal + w_obs*obstacle + w_vel*velocity (Q10)."""
    if lin is None:
        lin = linearize(p, Q, need_jac=False)
    return float(np.sum(lin.cost_pts)) + velocity_cost(p, Q[:, p.table.opt_qidx])


# --------------------------------------------------------------------------------------
# A13 replacement: projected Levenberg-Marquardt on the reduced problem (same algorithm as
# the CUDA solver, float64).  Unknowns: optimised rows of knots 2..T-1.
# --------------------------------------------------------------------------------------
@dataclass
class SolverOptions:
    max_iter: int = 100  # reference max_iter (gto_planner.py:141)
    tol_step: float = 1e-6  # |dq|_inf of an accepted step
    tol_grad: float = 1e-6  # |projected gradient|_inf
    lambda0: float = 1e-3
    lambda_min: float = 1e-9
    lambda_max: float = 1e9
    eta: float = 1e-4  # acceptance ratio
    noise_rel: float = 1e-6  # reductions below noise_rel * point-cost are inside fp32 noise
    bound_eps: float = 1e-12
    ftol: float = 1e-6  # STATUS_SLOW: an accepted step reduced the cost by <= ftol*f ...
    lambda_slow: float = 1e30  # ... while the damping that produced it was >= lambda_slow (1e30: test off)
    slow_window: int = 0  # >0: STATUS_SLOW as well when the cost fell by <= slow_ftol*f over the last slow_window iterations (off)
    slow_ftol: float = 1e-3
    as_rounds: int = 1  # active-set rounds per step: variables pushed beyond a limit are put on it, the rest re-solved
    lambda_reject: float = 1e-4  # a rejected step raises the damping to at least this value
    lambda_conv: float = 1e-2  # |dq| <= tol_step counts as converged only if the damping that produced the step was <= this
    bundle_radius: float = 3e-3  # pieces further than this (|.|_inf, rad) from the standing point are ignored (replaced first)
    bundle: int = 3  # cutting planes kept from points evaluated but not stood on (0: plain LM); see include/gto_b200.h gto_options.bundle


STATUS_CONVERGED = 0
STATUS_MAX_ITER = 1
STATUS_NAN = 2
STATUS_STALLED = 3
STATUS_SLOW = 4


def _system(p: Problem, Qx: np.ndarray, lin: Linearization):
    """Block-tridiagonal Gauss-Newton system for the free knots (2..T-1).
    Returns diag blocks D [m,n,n], constant off-diagonal scalar ``-a2`` (block (i,i+1) = -a2*I),
    and the gradient/2 ``gt`` [m,n]."""
    T, n = p.T, p.table.nopt
    a2 = p.w_vel / (p.dt * p.dt)
    m = T - 2
    D = np.zeros((m, n, n))
    gt = np.zeros((m, n))
    for i in range(m):
        t = i + 2
        cnt = 2.0 if t < T - 1 else 1.0
        D[i] = lin.H[t] + a2 * cnt * np.eye(n)
        gv = Qx[t] - Qx[t - 1]
        if t < T - 1:
            gv = gv - (Qx[t + 1] - Qx[t])
        gt[i] = lin.g[t] + a2 * gv
    return D, a2, gt


def _solve_block_tridiag(D: np.ndarray, off: float, rhs: np.ndarray) -> np.ndarray:
    """Block Cholesky (Thomas) for SPD block-tridiagonal [D_i, off*I]: returns x with A x = rhs."""
    m, n, _ = D.shape
    L = np.zeros_like(D)
    W = np.zeros_like(D)  # W_i = off * L_{i-1}^{-T}  -> sub-diagonal block of the factor
    y = np.zeros_like(rhs)
    for i in range(m):
        S = D[i].copy()
        if i > 0:
            S -= W[i] @ W[i].T
        L[i] = np.linalg.cholesky(S)
        b = rhs[i].copy()
        if i > 0:
            b -= W[i] @ y[i - 1]
        y[i] = np.linalg.solve(L[i], b)
        if i + 1 < m:
            # factor sub-diagonal block: A_{i+1,i} = off*I = W_{i+1} L_i^T  -> W_{i+1} = off * L_i^{-T}
            W[i + 1] = off * np.linalg.inv(L[i]).T
    x = np.zeros_like(rhs)
    for i in range(m - 1, -1, -1):
        b = y[i].copy()
        if i + 1 < m:
            b -= W[i + 1].T @ x[i + 1]
        x[i] = np.linalg.solve(L[i].T, b)
    return x


def _matvec(D, off, x):
    y = np.einsum("ijk,ik->ij", D, x)
    y[:-1] += off * x[1:]
    y[1:] += off * x[:-1]
    return y


BUNDLE_MAX = 4


def _total_gradient(p: Problem, Qx: np.ndarray, lin: Linearization) -> np.ndarray:
    """Half gradient of f over the free knots: J^T r of the point rows + the analytic velocity term."""
    return _system(p, Qx, lin)[2]


def _bundle_dual(bq: np.ndarray, M: np.ndarray) -> np.ndarray:
    """Dual of the bundle model: maximise ``b.theta + theta' M theta / 2`` over ``theta_1..K >= 0, sum <= 1`` (theta_0 = 1 - sum is
    the weight of the model at the standing point; row / column 0 of M and b_0 are zero) by pairwise exchange.  Same loop as
    bundle_dual() in gto_oracle.c and step_cr.cuh."""
    K = len(bq) - 1
    theta = np.zeros(K + 1)
    theta[0] = 1.0
    # the dual gradients of the pieces with weight all vanish at an interior optimum: the tolerance is relative to the largest |b_k|
    scale = float(np.max(np.abs(bq[1:]))) if K > 0 else 0.0
    for _ in range(12):
        G = bq + M[:, 1:] @ theta[1:]
        ib = 0
        for k in range(1, K + 1):
            if G[k] > G[ib]:
                ib = k
        jb = -1
        for k in range(K + 1):
            if theta[k] > 0.0 and (jb < 0 or G[k] < G[jb]):
                jb = k
        if jb < 0 or ib == jb or G[ib] - G[jb] <= 1e-13 * scale:
            break
        curv = -(M[ib, ib] - 2.0 * M[ib, jb] + M[jb, jb])
        delta = (G[ib] - G[jb]) / curv if curv > 0.0 else theta[jb]
        delta = min(delta, theta[jb])
        theta[ib] += delta
        theta[jb] -= delta
        if theta[jb] < 1e-15:
            theta[jb] = 0.0
    return theta


def lm_step(p: Problem, Qx: np.ndarray, lin: Linearization, lam: float, opts: SolverOptions, bundle=None):
    """One damped projected Gauss-Newton step of the bundle model.  ``bundle``: list of ``(g_k [m,n], e_k)`` cutting planes
    ``piece_k(s) = e_k + g_k.s`` (half-cost units relative to f/2, e_k <= 0).  Returns (trial Qx, d [m,n], pred, |proj grad|_inf);
    the trial is ``None`` when the projected gradient is already below ``tol_grad``."""
    tb = p.table
    D, a2, gt = _system(p, Qx, lin)
    X = Qx[2:]
    at_lo = (X <= tb.lo + opts.bound_eps) & (gt > 0)
    at_hi = (X >= tb.hi - opts.bound_eps) & (gt < 0)
    fixed = at_lo | at_hi
    pg = np.where(fixed, 0.0, gt)
    pgnorm = 2.0 * float(np.max(np.abs(pg))) if pg.size else 0.0
    if pgnorm <= opts.tol_grad:
        return None, np.zeros_like(gt), 0.0, pgnorm
    m, n = gt.shape
    bundle = bundle or []
    grads = [gt] + [gk for gk, _ in bundle]
    evals = [0.0] + [ek for _, ek in bundle]
    # Active-set rounds: a free variable that the step pushes beyond a joint limit is moved exactly onto the limit
    # (prescribed step dfix) and the other variables are re-solved with that step on the right-hand side -- one round of an
    # active-set method for the bound-constrained quadratic model.  Clipping alone distorts the coupled step: the model
    # then often predicts an increase and the step is rejected again and again while the damping rises.
    dfix = np.zeros_like(gt)
    as_round = 0
    while True:
        Dd = D.copy()
        for i in range(m):
            dg = np.diag(Dd[i]).copy()
            Dd[i] += lam * np.diag(dg)
            fi = fixed[i]
            if fi.any():
                Dd[i][fi, :] = 0.0
                Dd[i][:, fi] = 0.0
                Dd[i][fi, fi] = 1.0
        sols = []
        for gq in grads:
            rhs = np.zeros_like(gt)
            for i in range(m):
                fi = fixed[i]
                # free rows: -g - H[free, held] dfix[held] + a2 (dfix of the same joint at the neighbouring knots)
                Hoff = D[i] - np.diag(np.diag(D[i]))
                r = -gq[i] - Hoff @ (dfix[i] * fi)
                if i > 0:
                    r = r + a2 * dfix[i - 1] * fixed[i - 1]
                if i + 1 < m:
                    r = r + a2 * dfix[i + 1] * fixed[i + 1]
                rhs[i] = np.where(fi, dfix[i], r)
            sols.append(_solve_masked(Dd, -a2, rhs, fixed))
        K = len(bundle)
        theta = np.zeros(K + 1)
        theta[0] = 1.0
        if K > 0:
            bq = np.zeros(K + 1)
            M = np.zeros((K + 1, K + 1))
            for k in range(1, K + 1):
                dgk = grads[k] - gt
                bq[k] = evals[k] + float(np.sum(dgk * sols[0]))
                for j in range(1, K + 1):
                    M[k, j] = float(np.sum(dgk * (sols[j] - sols[0])))
            M = 0.5 * (M + M.T)
            theta = _bundle_dual(bq, M)
        d = sols[0].copy()
        for k in range(1, K + 1):
            if theta[k] != 0.0:
                d += theta[k] * (sols[k] - sols[0])
        if as_round >= opts.as_rounds:
            break
        Xn = X + d
        viol_lo = (~fixed) & (Xn < tb.lo)
        viol_hi = (~fixed) & (Xn > tb.hi)
        if not (viol_lo.any() or viol_hi.any()):
            break
        dfix = np.where(viol_lo, tb.lo - X, np.where(viol_hi, tb.hi - X, dfix))
        fixed = fixed | viol_lo | viol_hi
        as_round += 1
    Xn = np.clip(X + d, tb.lo, tb.hi)
    d = Xn - X
    Ad = _matvec(D, -a2, d)
    lin_term = float(np.sum(gt * d))
    for k in range(1, len(grads)):
        lin_term = max(lin_term, evals[k] + float(np.sum(grads[k] * d)))
    pred = -(lin_term + 0.5 * float(np.sum(d * Ad)))
    Qn = Qx.copy()
    Qn[2:] = Xn
    return Qn, d, pred, pgnorm


def _solve_masked(D, off, rhs, fixed):
    """Block-tridiagonal solve where the coupling block between knots i and i+1 is
    ``off * diag(free_i) diag(free_{i+1})`` restricted to matching joints (the coupling is a scalar
    times identity, so joint k of knot i only couples to joint k of knot i+1)."""
    m, n, _ = D.shape
    free = (~fixed).astype(np.float64)
    L = np.zeros_like(D)
    W = np.zeros_like(D)
    y = np.zeros_like(rhs)
    for i in range(m):
        S = D[i].copy()
        if i > 0:
            S -= W[i] @ W[i].T
        L[i] = np.linalg.cholesky(S)
        b = rhs[i].copy()
        if i > 0:
            b -= W[i] @ y[i - 1]
        y[i] = np.linalg.solve(L[i], b)
        if i + 1 < m:
            C = off * np.diag(free[i + 1] * free[i])  # A_{i+1,i}
            W[i + 1] = np.linalg.solve(L[i], C.T).T  # W L^T = C
    x = np.zeros_like(rhs)
    for i in range(m - 1, -1, -1):
        b = y[i].copy()
        if i + 1 < m:
            b -= W[i + 1].T @ x[i + 1]
        x[i] = np.linalg.solve(L[i].T, b)
    return x


@dataclass
class SolveResult:
    Q: np.ndarray  # [T,ndof]
    dQ: np.ndarray  # [T-1,ndof]
    cost: float
    iters: int
    status: int
    history: list


def unpack_solution(p: Problem, Q: np.ndarray):
    """A15 (optas/solver.py:126-159): full joint state with parameter rows re-inflated;
    ``dQ`` optimised rows = finite differences (exactly the eliminated equality constraints),
    parameter rows = 0 (``<robot>/dq/p`` defaults to zeros, optas/mx_container.py:121)."""
    tb = p.table
    dQ = np.zeros((p.T - 1, tb.ndof))
    dQ[:, tb.opt_qidx] = np.diff(Q[:, tb.opt_qidx], axis=0) / p.dt
    return Q, dQ


def _piece_value(gk, Fk, dyk, F, opts) -> float:
    """Value at the standing point (half-cost units, forced <= 0) of the cutting plane taken at ``x + dyk``; a plane further than
    ``bundle_radius`` away is switched off."""
    if float(np.max(np.abs(dyk))) > opts.bundle_radius:
        return -1e300
    return -abs(0.5 * (Fk - F) - float(np.sum(gk * dyk)))


def solve_lm(p: Problem, opts: Optional[SolverOptions] = None) -> SolveResult:
    opts = opts or SolverOptions()
    tb = p.table
    oi = tb.opt_qidx
    Q = initial_trajectory(p)
    lin = linearize(p, Q)
    F = float(np.sum(lin.cost_pts)) + velocity_cost(p, Q[:, oi])
    lam, nu = opts.lambda0, 2.0
    status = STATUS_MAX_ITER
    hist = [F]
    fhist = [0.0] * 16
    fhist[0] = F
    it = 0
    KB = min(max(int(opts.bundle), 0), BUNDLE_MAX)
    bun = []  # cutting planes [g_k, f_k, y_k - x] of points evaluated but not stood on (rejected trials, iterates left)
    while it < opts.max_iter:
        pieces = [(gk, _piece_value(gk, Fk, dyk, F, opts)) for gk, Fk, dyk in bun]
        Qx_trial, d, pred, pgnorm = lm_step(p, Q[:, oi], lin, lam, opts, pieces)
        if Qx_trial is None:
            status = STATUS_CONVERGED
            break
        it += 1
        Qt = Q.copy()
        Qt[:, oi] = Qx_trial
        lin_t = linearize(p, Qt)
        Fp_t = float(np.sum(lin_t.cost_pts))
        Ft = Fp_t + velocity_cost(p, Qx_trial)
        if not np.isfinite(Ft):
            status = STATUS_NAN
            break
        ared = 0.5 * (F - Ft)
        noise = opts.noise_rel * max(float(np.sum(lin.cost_pts)), Fp_t)
        step = float(np.max(np.abs(d))) if d.size else 0.0
        acc = pred > 0 and ared + noise >= opts.eta * pred
        if KB > 0:
            # bundle update: the point we do not stand on after this decision becomes a cutting plane; kept pieces are re-based
            # to the new standing point; a full bundle replaces its least active piece (most negative value there)
            Fnew = Ft if acc else F
            if acc:
                for piece in bun:
                    piece[2] = piece[2] - d
                new_piece = [_total_gradient(p, Q[:, oi], lin), F, -d]
            else:
                new_piece = [_total_gradient(p, Qx_trial, lin_t), Ft, d.copy()]
            if len(bun) < KB:
                bun.append(new_piece)
            else:
                ev = [_piece_value(gk, Fk, dyk, Fnew, opts) for gk, Fk, dyk in bun]
                slot = 0
                for k in range(1, len(bun)):
                    if ev[k] < ev[slot]:
                        slot = k
                bun[slot] = new_piece
        if acc:
            rho = ared / pred if pred > 0 else 1.0
            F_before, lam_used = F, lam
            Q, lin, F = Qt, lin_t, Ft
            # gain ratio clamped to [0, 1]: a step accepted only thanks to the noise allowance can have rho << 0, and
            # Nielsen's cubic would then multiply the damping by hundreds in one step
            lam = max(opts.lambda_min, lam * max(1.0 / 3.0, 1.0 - (2.0 * min(max(rho, 0.0), 1.0) - 1.0) ** 3))
            nu = 2.0
            hist.append(F)
            if step <= opts.tol_step:
                # a small step certifies a stationary point only when it was (nearly) the undamped Gauss-Newton step;
                # under heavy damping the iterate rests on a gradient jump of the trilinear field (not converged)
                status = STATUS_CONVERGED if lam_used <= opts.lambda_conv else STATUS_SLOW
                break
            if lam_used >= opts.lambda_slow and ared <= opts.ftol * F_before:
                status = STATUS_SLOW  # heavily damped and no longer reducing the cost: a kink of the trilinear field
                break
        else:
            if pred <= 0 and step <= opts.tol_step:
                status = STATUS_CONVERGED if lam <= opts.lambda_conv else STATUS_SLOW
                break
            lam = min(opts.lambda_max, max(lam * nu, opts.lambda_reject))
            nu *= 2.0
            if lam >= opts.lambda_max:
                status = STATUS_STALLED
                break
        if opts.slow_window > 0:  # windowed progress test on the accepted cost (same bookkeeping as k_step)
            if it >= opts.slow_window and fhist[(it - opts.slow_window) % 16] - F <= opts.slow_ftol * F:
                status = STATUS_SLOW
                break
            fhist[it % 16] = F
    Qf, dQ = unpack_solution(p, Q)
    return SolveResult(Qf, dQ, F, it, status, hist)


def solve_scipy(p: Problem, xtol: float = 1e-14, gtol: float = 1e-12, max_nfev: int = 400):
    """Independent cross-check: SciPy trust-region-reflective bound-constrained least squares on
    the same residual vector (reduced space)."""
    from scipy.optimize import least_squares

    tb = p.table
    oi = tb.opt_qidx
    n, T = tb.nopt, p.T
    Q0 = initial_trajectory(p)
    sv = np.sqrt(p.w_vel) / p.dt

    dense = (T - 2) * n <= 400

    def unpack(x):
        Q = Q0.copy()
        Q[2:, oi] = x.reshape(T - 2, n)
        return Q

    def fun(x):
        Q = unpack(x)
        lin = linearize(p, Q, need_jac=False)
        parts = [lin.r_obs.reshape(-1), lin.r_goal.reshape(-1), lin.r_stand.reshape(-1), (sv * np.diff(Q[:, oi], axis=0)).reshape(-1)]
        return np.concatenate(parts)

    def jac(x):
        Q = unpack(x)
        lin = linearize(p, Q)
        P = tb.npoints
        rows = T * P + lin.r_goal.size + lin.r_stand.size + (T - 1) * n
        from scipy.sparse import lil_matrix

        J = lil_matrix((rows, (T - 2) * n))
        for t in range(2, T):
            nz = np.nonzero(np.any(lin.J_obs[t] != 0, axis=1))[0]
            if nz.size:
                J[t * P + nz[:, None], (t - 2) * n + np.arange(n)[None, :]] = lin.J_obs[t][nz]
        base = T * P
        J[base : base + lin.r_goal.size, (T - 3) * n : (T - 2) * n] = lin.J_goal.reshape(-1, n)
        base += lin.r_goal.size
        if p.use_standoff:
            ks = p.knot_standoff
            J[base : base + lin.r_stand.size, (ks - 2) * n : (ks - 1) * n] = lin.J_stand.reshape(-1, n)
        base += lin.r_stand.size
        for t in range(T - 1):  # row block t: sv*(Q_{t+1}-Q_t)
            for k in range(n):
                if t + 1 >= 2:
                    J[base + t * n + k, (t + 1 - 2) * n + k] = sv
                if t >= 2:
                    J[base + t * n + k, (t - 2) * n + k] = -sv
        return J.toarray() if dense else J.tocsr()

    x0 = Q0[2:, oi].reshape(-1)
    lo = np.tile(tb.lo, T - 2)
    hi = np.tile(tb.hi, T - 2)
    x0 = np.clip(x0, lo + 1e-12, hi - 1e-12)
    res = least_squares(fun, x0, jac=jac, bounds=(lo, hi), method="trf", xtol=xtol, ftol=1e-15, gtol=gtol, max_nfev=max_nfev, tr_solver="exact" if dense else "lsmr", x_scale=1.0)
    # tr_solver exact needs dense J
    Q = unpack(res.x)
    return Q, 2.0 * res.cost, res


# --------------------------------------------------------------------------------------
# A14: seeds and seed ranking
# --------------------------------------------------------------------------------------
def interpolate_seed(qc: np.ndarray, q_goal: np.ndarray, T: int) -> np.ndarray:
    """Closed form of ``interpolate_waypoints(np.stack([qc, q_goal]), T, ndof)`` (gto/utils.py:63-82):
    a clamped cubic through two waypoints is the smoothstep, sampled at the interior of
    ``linspace(0,1,T+2)``.  Returns [T,ndof]."""
    s = (np.arange(T) + 1.0) / (T + 1.0)
    w = 3 * s**2 - 2 * s**3
    qc = np.asarray(qc, dtype=np.float64).reshape(1, -1)
    qg = np.asarray(q_goal, dtype=np.float64).reshape(1, -1)
    return qc + (qg - qc) * w[:, None]


def make_seed(table, qc, q_goal, T: int, interpolate: bool = True, standoff_offset: int = -10) -> np.ndarray:
    """Seed of ``plan`` / ``plan_goalset`` (gto_planner.py:150-158, 199-219): interpolated plan
    with parameter-joint rows overwritten by ``qc``; ``interpolate=False`` keeps ``qc`` and only sets
    the last ``|standoff_offset|`` knots to the final column of the interpolated plan."""
    qc = np.asarray(qc, dtype=np.float64).reshape(-1)
    plan = interpolate_seed(qc, q_goal, T)
    plan[:, table.par_qidx] = qc[table.par_qidx]
    if interpolate:
        return plan
    Q0 = np.tile(qc, (T, 1))
    Q0[T + standoff_offset :] = plan[T - 1]
    return Q0


def plan_cost_nearest(table, plan: np.ndarray, fld: Field, base_position) -> tuple:
    """``compute_plan_cost`` (gto/gto_models.py:204-215): sum of nearest-node costs over all knots
    (clip-then-truncate indexing of ``points_to_offsets_numpy``) and ``|q_0 - q_{T-1}|``.
    ``plan`` is [T,ndof]."""
    cost = 0.0
    flat = fld.cost.reshape(-1)
    for t in range(plan.shape[0]):
        W = world_points(table, plan[t], base_position)
        cost += float(np.sum(flat[fld.offsets_numpy_twin(W)]))
    return cost, float(np.linalg.norm(plan[0] - plan[-1]))
