"""Synthetic scenes, goals and seeds for the BASELINE configs.

A scene is a union of axis-aligned boxes; its exact signed distance is evaluated at the field nodes
``origin + k*pitch`` and pushed through the reference's cost transform
(``mesh_to_sdf/depth_point_cloud.py:84-89``, epsilon = 0.02 m) to float32 -- the same kind of field the
reference builds from a depth image and hands to the planner as ``sdf_cost_all`` / ``sdf_cost_obstacle``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .kinematics import ee_frames, gripper_frames
from .robot_table import RobotTable


@dataclass
class CostField:
    """Node-centred voxel cost field (C-order ``[Nx,Ny,Nz]``, nodes at ``origin + k*pitch``)."""

    cost: np.ndarray  # float32 [Nx,Ny,Nz]
    origin: np.ndarray  # [3]
    pitch: float


def box_sdf(points: np.ndarray, boxes: Sequence[Tuple[Sequence[float], Sequence[float]]]) -> np.ndarray:
    """Exact signed distance to a union of axis-aligned boxes given as (center, full size)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    d = np.full(pts.shape[0], np.inf)
    for center, size in boxes:
        q = np.abs(pts - np.asarray(center, dtype=np.float64)) - 0.5 * np.asarray(size, dtype=np.float64)
        outside = np.linalg.norm(np.maximum(q, 0.0), axis=1)
        inside = np.minimum(np.max(q, axis=1), 0.0)
        d = np.minimum(d, outside + inside)
    return d


def sdf_cost(distances: np.ndarray, epsilon: float = 0.02, w_inside: float = 1.0) -> np.ndarray:
    d = np.asarray(distances, dtype=np.float32)
    cost = np.zeros_like(d)
    inside = d < 0
    cost[inside] = w_inside * (-d[inside] + epsilon / 2)
    shell = (d > 0) & (d < epsilon)
    cost[shell] = np.square(d[shell] - epsilon) / (2 * epsilon)
    return cost


def make_field(boxes, lo, hi, n: int, margin: float = 0.4, epsilon: float = 0.02) -> CostField:
    """Field of n^3 nodes covering [lo - margin, hi + margin] (``origin = workspace min - 0.4``,
    reference ``gto/gto_models.py:140``); pitch = largest extent / (n - 1).

    The exact signed distance to the union of boxes is the minimum over the boxes; the cost is zero wherever that distance
    is >= epsilon, so every box only needs to be evaluated on the sub-grid within epsilon (plus one node) of it -- the result is
    identical to evaluating every box at every node."""
    lo = np.asarray(lo, dtype=np.float64) - margin
    hi = np.asarray(hi, dtype=np.float64) + margin
    pitch = float(np.max(hi - lo) / (n - 1))
    ax = [lo[a] + pitch * np.arange(n) for a in range(3)]
    dmin = np.full((n, n, n), np.inf)
    for center, size in boxes:
        c, h = np.asarray(center, dtype=np.float64), 0.5 * np.asarray(size, dtype=np.float64)
        sl, q = [], []
        for a in range(3):
            i0 = int(np.searchsorted(ax[a], c[a] - h[a] - epsilon - pitch, side="left"))
            i1 = int(np.searchsorted(ax[a], c[a] + h[a] + epsilon + pitch, side="right"))
            sl.append(slice(max(i0 - 1, 0), min(i1 + 1, n)))
            q.append(np.abs(ax[a][sl[a]] - c[a]) - h[a])
        if any(s_.stop <= s_.start for s_ in sl):
            continue
        qx, qy, qz = q[0][:, None, None], q[1][None, :, None], q[2][None, None, :]
        # same expression as box_sdf: |max(q, 0)|_2 + min(max(q), 0)
        outside = np.sqrt(np.maximum(qx, 0.0) ** 2 + np.maximum(qy, 0.0) ** 2 + np.maximum(qz, 0.0) ** 2)
        inside = np.minimum(np.maximum(np.maximum(qx, qy), qz), 0.0)
        sub = dmin[sl[0], sl[1], sl[2]]
        np.minimum(sub, outside + inside, out=sub)
    cost = sdf_cost(dmin.reshape(-1), epsilon).reshape(n, n, n) if len(boxes) else np.zeros((n, n, n), dtype=np.float32)
    return CostField(np.ascontiguousarray(cost, dtype=np.float32), lo, pitch)


def _make_field_dense(boxes, lo, hi, n: int, margin: float = 0.4, epsilon: float = 0.02) -> CostField:
    """Every box at every node (the definition; kept to test make_field against)."""
    lo = np.asarray(lo, dtype=np.float64) - margin
    hi = np.asarray(hi, dtype=np.float64) + margin
    pitch = float(np.max(hi - lo) / (n - 1))
    ax = [lo[a] + pitch * np.arange(n) for a in range(3)]
    cost = np.zeros((n, n, n), dtype=np.float32)
    if len(boxes):
        yz = np.stack(np.meshgrid(ax[1], ax[2], indexing="ij"), axis=-1).reshape(-1, 2)
        for i, x in enumerate(ax[0]):
            pts = np.concatenate([np.full((yz.shape[0], 1), x), yz], axis=1)
            cost[i] = sdf_cost(box_sdf(pts, boxes), epsilon).reshape(n, n)
    return CostField(cost, lo, pitch)


def workspace_box(arm_len: float, arm_height: float):
    """``setup_workspace_field`` extents (reference ``gto/gto_models.py:135-139``)."""
    return np.array([0.0, -arm_len, 0.0]), np.array([arm_len, arm_len, arm_height + arm_len])


def tabletop_boxes(z_top: float = 0.0, thickness: float = 0.05, x=(0.1, 1.3), y=(-0.8, 0.8)):
    return [((0.5 * (x[0] + x[1]), 0.5 * (y[0] + y[1]), z_top - 0.5 * thickness), (x[1] - x[0], y[1] - y[0], thickness))]


def shelf_boxes(x0: float = 0.9, z0: float = 0.0):
    """Six panels with the dimensions of the reference's shelf prop (``data/objects/shelf/shelf.urdf``),
    placed at (x0, 0, z0) and yawed 180 degrees (``examples/pybullet_scenereplica.py:218-226``)."""
    return [
        ((x0, 0.0, z0 + 0.0), (0.392, 0.803, 0.022)),
        ((x0, 0.0, z0 + 0.781), (0.392, 0.803, 0.022)),
        ((x0, 0.3925, z0 + 0.3905), (0.392, 0.018, 0.759)),
        ((x0, -0.3925, z0 + 0.3905), (0.392, 0.018, 0.759)),
        ((x0, 0.0, z0 + 0.3905), (0.374, 0.767, 0.022)),
        ((x0 + 0.187, 0.0, z0 + 0.3905), (0.018, 0.767, 0.759)),
    ]


def clutter_boxes(rng: np.random.Generator, n: int = 20, x=(0.25, 0.9), y=(-0.6, 0.6)):
    boxes = tabletop_boxes()
    for _ in range(n):
        size = rng.uniform(0.05, 0.20, size=3)
        c = np.array([rng.uniform(*x), rng.uniform(*y), 0.5 * size[2]])
        boxes.append((tuple(c), tuple(size)))
    return boxes


def smoothstep_seed(qc: np.ndarray, q_goal: np.ndarray, T: int) -> np.ndarray:
    """Closed form of the reference seed (clamped cubic through two waypoints sampled at the interior of
    ``linspace(0,1,T+2)``; ``gto/utils.py:63-82``).  qc [ndof], q_goal [B,ndof] -> [B,T,ndof]."""
    s = (np.arange(T) + 1.0) / (T + 1.0)
    w = (3 * s**2 - 2 * s**3)[None, :, None]
    qc = np.asarray(qc, dtype=np.float64).reshape(1, 1, -1)
    qg = np.asarray(q_goal, dtype=np.float64)
    qg = qg.reshape(-1, 1, qg.shape[-1])
    return qc + (qg - qc) * w


def sample_goals(table: RobotTable, qc: np.ndarray, n: int, rng: np.random.Generator, lo_ws, hi_ws,
                 boxes=(), clearance: float = 0.05, seed_noise: float = 0.05, max_tries: int = 200):
    """Reachable-by-construction goals: ``RT = FK_ee(q*)`` for ``q*`` uniform within the joint limits,
    kept when the gripper origin lies in the workspace box and >= ``clearance`` away from every scene box.
    Returns (RT [n,4,4], q_star [n,ndof], q_seed_goal [n,ndof] = q* + N(0, seed_noise) clipped)."""
    qc = np.asarray(qc, dtype=np.float64)
    RT, QS = [], []
    tries = 0
    while len(RT) < n:
        tries += 1
        if tries > max_tries:
            raise RuntimeError("could not sample enough reachable goals")
        m = max(4 * (n - len(RT)), 64)
        q = np.tile(qc, (m, 1))
        q[:, table.opt_qidx] = rng.uniform(table.lo, table.hi, size=(m, table.nopt))
        Tg = gripper_frames(table, q)
        pos = Tg[:, :3, 3]
        ok = np.all((pos >= lo_ws) & (pos <= hi_ws), axis=1)
        if len(boxes):
            ok &= box_sdf(pos, boxes) >= clearance
        Te = ee_frames(table, q[ok])
        for i in range(Te.shape[0]):
            if len(RT) < n:
                RT.append(Te[i])
                QS.append(q[ok][i])
    RT = np.stack(RT)
    QS = np.stack(QS)
    QG = QS.copy()
    QG[:, table.opt_qidx] = np.clip(QS[:, table.opt_qidx] + rng.normal(0.0, seed_noise, size=(n, table.nopt)), table.lo, table.hi)
    return RT, QS, QG


def sample_grasps_around(table: RobotTable, qc: np.ndarray, center, n: int, rng: np.random.Generator, approach_axis: str = "z",
                         reach: float = 0.10, pos_tol: float = 0.03, approach_dir=(0.0, 0.0, -1.0), min_cos: float = 0.3,
                         boxes=(), clearance: float = 0.02, seed_noise: float = 0.05, max_rounds: int = 2000, batch: int = 8192):
    """Candidate grasps of ONE object (BASELINE "candidate grasps"): configurations q* whose grasp point -- ``reach``
    metres ahead of the gripper frame along its approach axis -- lies within ``pos_tol`` of ``center`` and whose approach
    axis points roughly along ``approach_dir``.  Reachable by construction: all optimised joints but the first revolute
    one are drawn uniformly within their limits, and that joint is then solved in closed form so that the grasp point
    swings onto the object's azimuth.  Returns (RT [n,4,4] pose of link_ee, q_star [n,ndof], q_goal_seed [n,ndof])."""
    from .kinematics import fk_movable, gripper_frames, ee_frames
    from .robot_table import JOINT_REVOLUTE

    qc = np.asarray(qc, dtype=np.float64)
    center = np.asarray(center, dtype=np.float64)
    adir = np.asarray(approach_dir, dtype=np.float64)
    adir = adir / np.linalg.norm(adir)
    ax = "xyz".index(approach_axis)
    j0 = next(j for j in range(table.nmov) if table.mov_opt[j] >= 0 and table.mov_type[j] == JOINT_REVOLUTE)
    k0 = int(table.mov_opt[j0])
    keepQ = []
    for _ in range(max_rounds):
        q = np.tile(qc, (batch, 1))
        q[:, table.opt_qidx] = rng.uniform(table.lo, table.hi, size=(batch, table.nopt))
        q[:, table.opt_qidx[k0]] = 0.0
        Tm = fk_movable(table, q)
        o = Tm[:, j0, :3, 3]
        z = Tm[:, j0, :3, :3] @ table.mov_axis[j0]
        Tg = gripper_frames(table, q)
        p = Tg[:, :3, 3] + reach * Tg[:, :3, ax]

        def cyl(v):
            d = v - o
            h = np.sum(d * z, axis=1)
            rv = d - h[:, None] * z
            return h, rv

        hp, rp = cyl(p)
        hc, rc = cyl(center[None, :])
        rpn, rcn = np.linalg.norm(rp, axis=1), np.linalg.norm(rc, axis=1)
        ok = (np.abs(hp - hc) < pos_tol) & (np.abs(rpn - rcn) < pos_tol) & (rpn > 1e-3)
        theta = np.arctan2(np.sum(np.cross(rp, rc) * z, axis=1), np.sum(rp * rc, axis=1))
        ok &= (theta >= table.lo[k0]) & (theta <= table.hi[k0])
        q = q[ok]
        q[:, table.opt_qidx[k0]] = theta[ok]
        if q.shape[0] == 0:
            continue
        Tg = gripper_frames(table, q)
        good = (Tg[:, :3, ax] @ adir) > min_cos
        if len(boxes):
            good &= box_sdf(Tg[:, :3, 3], boxes) >= clearance
        keepQ.append(q[good])
        if sum(x.shape[0] for x in keepQ) >= n:
            break
    QS = np.concatenate(keepQ)[:n] if keepQ else np.zeros((0, table.ndof))
    if QS.shape[0] < n:
        raise RuntimeError(f"only {QS.shape[0]} of {n} candidate grasps found")
    RT = ee_frames(table, QS)
    QG = QS.copy()
    QG[:, table.opt_qidx] = np.clip(QS[:, table.opt_qidx] + rng.normal(0.0, seed_noise, size=(n, table.nopt)), table.lo, table.hi)
    return RT, QS, QG
