"""Host-side helpers with the names of the reference's ``gto/utils.py`` (hot-path subset).

``interpolate_waypoints`` is the seed generator of row A14 (reference ``gto/utils.py:63-82``);
the VTK/toppra debugging helpers of the reference are rendering-only and not reproduced."""
from __future__ import annotations

import os

import numpy as np
import yaml
from scipy import interpolate


def get_root_dir():
    """Root under which ``data/configs`` and ``data/robots`` live.  Defaults to the directory that
    contains this package's ``compat`` folder; override with ``GTO_ROOT_DIR`` to point at a
    GraspTrajOpt checkout."""
    env = os.environ.get("GTO_ROOT_DIR")
    if env:
        return env
    return os.path.join(os.path.dirname(__file__), "..")


def load_yaml(file_path):
    if isinstance(file_path, str):
        with open(file_path) as fh:
            return yaml.load(fh, Loader=yaml.Loader)
    return file_path


def rotZ(rotz):
    c, s = np.cos(rotz), np.sin(rotz)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)


def interpolate_waypoints(waypoints, n, m, mode="cubic"):
    """Waypoints [k, m] -> [n, m] samples at the *interior* of ``linspace(0, 1, n+2)``; "cubic" is a
    clamped cubic spline (zero end slopes), "linear" piecewise linear."""
    waypoints = np.asarray(waypoints, dtype=np.float64)
    data = np.zeros([n, m])
    x = np.linspace(0, 1, waypoints.shape[0])
    t = np.linspace(0, 1, n + 2)[1:-1]
    for i in range(waypoints.shape[1]):
        y = waypoints[:, i]
        if mode == "linear":
            f = interpolate.interp1d(x, y, "linear")
        elif mode == "cubic":
            f = interpolate.CubicSpline(x, y, bc_type="clamped")
        else:
            raise ValueError(f"unknown interpolation mode {mode}")
        data[:, i] = f(t)
    return data


def default_pose(robot_model):
    """Default joint command per robot (reference ``gto/utils.py:36-60``)."""
    q = np.zeros((robot_model.ndof,), dtype=np.float32)
    if robot_model.name == "fetch":
        q[2] = 0.4
        q[3] = 0.009195
        q[4] = 0.908270
        q[[5, 6, 7, 8, 9, 10, 11]] = [1.32, 0.7, 0.0, -2.0, 0.0, -0.57, 0.0]
    elif robot_model.name == "panda":
        q = np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785, 0.0, 0.0])
    return q


def plan_collision_audit(robot, plan, depth_pc, base_position=None, min_points=5):
    """Collision audit of a stored plan (reference ``examples/pybullet_evaluate_plans.py:219-237``): a knot is in collision when
    more than ``min_points`` robot surface points lie behind the observed surface (``DepthPointCloud.get_sdf < 0``).

    The reference queries the KD-tree once per knot and stops at the first hit; here the surface points of ALL knots go through
    one ``get_sdf`` call (one ``k_cloud_query`` launch with the GPU backend).  ``plan`` is ndof-by-T.
    Returns ``(in_collision, first_knot or -1, counts [T])``."""
    plan = np.asarray(plan, dtype=np.float64)
    T = plan.shape[1]
    off = np.zeros((1, 3)) if base_position is None else np.asarray(base_position, dtype=np.float64).reshape(1, 3)
    pts = np.concatenate([robot.compute_fk_surface_points(plan[:, i])[0] + off for i in range(T)])
    sdf = np.asarray(depth_pc.get_sdf(pts)).reshape(T, -1)
    counts = (sdf < 0).sum(axis=1)
    hit = np.flatnonzero(counts > min_points)
    return bool(hit.size), (int(hit[0]) if hit.size else -1), counts
