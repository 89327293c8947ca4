"""Batched NumPy forward kinematics on a ``RobotTable`` (host-side input preparation only:
synthetic goal poses, seeds, seed ranking).  The per-iteration FK of the solver runs in CUDA."""
from __future__ import annotations

import numpy as np

from .robot_table import RobotTable, JOINT_REVOLUTE


def _hom(T34):
    T = np.zeros(T34.shape[:-2] + (4, 4))
    T[..., :3, :] = T34
    T[..., 3, 3] = 1.0
    return T


def fk_movable(table: RobotTable, Q: np.ndarray) -> np.ndarray:
    """Q [B,ndof] -> frames of all movable joints after their motion, [B,nmov,4,4]."""
    Q = np.asarray(Q, dtype=np.float64).reshape(-1, table.ndof)
    B = Q.shape[0]
    out = np.zeros((B, table.nmov, 4, 4))
    for j in range(table.nmov):
        P = np.broadcast_to(np.eye(4), (B, 4, 4)) if table.mov_parent[j] < 0 else out[:, table.mov_parent[j]]
        O = _hom(table.mov_origin[j])
        M = np.broadcast_to(np.eye(4), (B, 4, 4)).copy()
        q = Q[:, table.mov_qidx[j]]
        a = table.mov_axis[j]
        if table.mov_type[j] == JOINT_REVOLUTE:
            K = np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])
            M[:, :3, :3] = np.eye(3) + np.sin(q)[:, None, None] * K + (1.0 - np.cos(q))[:, None, None] * (K @ K)
        else:
            M[:, :3, 3] = q[:, None] * a[None, :]
        out[:, j] = P @ O @ M
    return out


def gripper_frames(table: RobotTable, Q: np.ndarray) -> np.ndarray:
    """Plain link frame of ``link_gripper`` for every configuration, [B,4,4]."""
    Tm = fk_movable(table, Q)
    P = np.broadcast_to(np.eye(4), (Tm.shape[0], 4, 4)) if table.grip_mov < 0 else Tm[:, table.grip_mov]
    return P @ _hom(table.grip_tf)


def ee_frames(table: RobotTable, Q: np.ndarray) -> np.ndarray:
    """Frame of ``link_ee``: gripper frame x inverse(G)."""
    G = _hom(table.G)
    return gripper_frames(table, Q) @ np.linalg.inv(G)


def world_points(table: RobotTable, Q: np.ndarray, base_position=None) -> np.ndarray:
    """All surface points for every configuration, [B,P,3] (robot base frame + base_position)."""
    Tm = fk_movable(table, Q)
    B = Tm.shape[0]
    W = np.zeros((B, table.npoints, 3))
    for l in range(table.nlinks):
        P = np.broadcast_to(np.eye(4), (B, 4, 4)) if table.link_mov[l] < 0 else Tm[:, table.link_mov[l]]
        F = P @ _hom(table.link_tf[l])
        s, c = int(table.link_pt_start[l]), int(table.link_pt_count[l])
        W[:, s : s + c] = np.einsum("bij,pj->bpi", F[:, :3, :3], table.points[s : s + c]) + F[:, None, :3, 3]
    if base_position is not None:
        W = W + np.asarray(base_position, dtype=np.float64).reshape(-1, 1, 3)
    return W
