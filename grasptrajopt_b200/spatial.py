"""NumPy float64 spatial-math primitives with the reference's conventions.

Conventions follow ``optas/spatialmath.py``: ``rpy2r`` default order "zyx" is
``Rz(yaw) @ Ry(pitch) @ Rx(roll)`` with ``rpy = (roll, pitch, yaw)`` (:187-211);
``angvec2r`` is Rodrigues' formula about the normalised axis (:91-100);
``invt`` inverts a homogeneous transform (:261-280); ``standoff`` is a pure
translation along a named axis (:160-183).
"""
from __future__ import annotations

import numpy as np

pi = np.pi
eps = np.finfo(float).eps


def rotx(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def roty(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def rotz(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rpy2r(rpy, opt: str = "zyx"):
    r, p, y = (float(v) for v in np.asarray(rpy, dtype=np.float64).reshape(-1))
    if opt in ("xyz", "arm"):
        return rotx(y) @ roty(p) @ rotz(r)
    if opt in ("zyx", "vehicle"):
        return rotz(y) @ roty(p) @ rotx(r)
    if opt in ("yxz", "camera"):
        return roty(y) @ rotx(p) @ rotz(r)
    raise ValueError(f"didn't recognize given option {opt}")


def skew(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    if v.shape[0] == 1:
        return np.array([[0.0, -v[0]], [v[0], 0.0]])
    if v.shape[0] == 3:
        return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    raise ValueError("expecting a scalar or 3-vector")


def unit(v):
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    return v / np.linalg.norm(v)


def angvec2r(theta, v):
    sk = skew(unit(v))
    return np.eye(3) + np.sin(theta) * sk + (1.0 - np.cos(theta)) * (sk @ sk)


def rt2tr(R, t):
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
    return T


def r2t(R):
    return rt2tr(R, np.zeros(3))


def t2r(T):
    return np.asarray(T)[:3, :3]


def transl(T):
    return np.asarray(T)[:3, 3]


def invt(T):
    T = np.asarray(T, dtype=np.float64)
    R = T[:3, :3]
    return rt2tr(R.T, -R.T @ T[:3, 3])


def standoff(offset: float, axis: str = "x"):
    T = np.eye(4)
    T["xyz".index(axis), 3] = float(offset)
    return T


def mat2quat_wxyz(R):
    """Rotation matrix -> unit quaternion (w,x,y,z), w >= 0 (transforms3d convention used at
    ``gto/ik_solver.py:95-97``)."""
    R = np.asarray(R, dtype=np.float64)
    K = np.array(
        [
            [R[0, 0] - R[1, 1] - R[2, 2], 0.0, 0.0, 0.0],
            [R[0, 1] + R[1, 0], R[1, 1] - R[0, 0] - R[2, 2], 0.0, 0.0],
            [R[0, 2] + R[2, 0], R[1, 2] + R[2, 1], R[2, 2] - R[0, 0] - R[1, 1], 0.0],
            [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]],
        ]
    ) / 3.0
    w, v = np.linalg.eigh(K)
    q = v[[3, 0, 1, 2], np.argmax(w)]
    return -q if q[0] < 0 else q
