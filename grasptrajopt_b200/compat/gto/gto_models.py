"""``GTORobotModel``: robot = URDF kinematics + per-link surface point sets + voxel-field geometry.

Host-side mirror of the reference class (``gto/gto_models.py:23-292``) without CasADi/trimesh:
same constructor arguments, attributes (``surface_pc_map``, ``visual_tf``, ``field_margin``,
``grid_resolution``, ``origin``, ``field_shape``, ``field_size``, ``workspace_points``) and methods.
New, B200-specific parts: ``sample_point_count``/``seed`` (the reference hard-codes 100 unseeded
samples per link, :76-77) and ``to_table()`` which flattens the model for the CUDA kernels.
"""
from __future__ import annotations

import os
import zlib
from typing import List, Optional

import numpy as np
from sklearn.neighbors import KDTree

import optas
from optas.models import RobotModel
from optas.dm import DM, _as2d
from grasptrajopt_b200 import spatial as sp
from grasptrajopt_b200.meshio import load_mesh, sample_surface, SurfacePointCloud
from grasptrajopt_b200.robot_table import build_robot_table, RobotTable


class _VisualTf:
    """``visual_tf[name](q)`` -> object with ``.toarray()`` (reference returns a CasADi Function,
    ``gto_models.py:100,108``)."""

    def __init__(self, model: "GTORobotModel", link: str, vis: np.ndarray):
        self._model, self._link, self._vis = model, link, vis

    def __call__(self, q) -> DM:
        q = _as2d(q).reshape(-1)
        return DM(self._model._global_tf(self._link, q) @ self._vis)


class GTORobotModel(RobotModel):
    def __init__(
        self,
        model_dir,
        urdf_filename: Optional[str] = None,
        urdf_string: Optional[str] = None,
        xacro_filename: Optional[str] = None,
        name: Optional[str] = None,
        time_derivs: List[int] = [0],
        qddlim=None,
        T: Optional[int] = None,
        param_joints: List[str] = [],
        collision_link_names=None,
        sample_point_count: int = 100,
        seed: int = 0,
        surface_pc_map=None,
    ):
        super().__init__(urdf_filename, urdf_string, xacro_filename, name, time_derivs, qddlim, T, param_joints)
        self.model_dir = model_dir
        self.collision_link_names = collision_link_names
        self.sample_point_count = sample_point_count
        self.seed = seed
        self.surface_pc_map = surface_pc_map if surface_pc_map is not None else self.compute_link_surface_points()
        self.visual_tf = self.setup_fk_functions()
        self.field_margin = 0.4
        self.grid_resolution = 0.05
        self._tables = {}

    def get_standoff_pose(self, offset, axis):
        pose = np.eye(4, dtype=np.float32)
        if axis in "xyz" and len(axis) == 1:
            pose["xyz".index(axis), 3] = offset
        else:
            print("unknow standoff axis", axis)
        return pose

    # -- A3: per-link surface point sets ------------------------------------------------------
    def _collision_links(self):
        for link in self.urdf.links:
            if link.visual is None or link.visual.geometry is None:
                continue
            if self.collision_link_names is None or link.name in self.collision_link_names:
                yield link

    def compute_link_surface_points(self):
        out = {}
        for link in self._collision_links():
            filename = os.path.join(self.model_dir, link.visual.geometry.filename)
            mesh = load_mesh(filename, scale=link.visual.geometry.scale)
            # one independent, reproducible stream per link
            rng = np.random.default_rng([int(self.seed), zlib.crc32(link.name.encode())])
            pts, nrm = sample_surface(mesh, self.sample_point_count, rng)
            out[link.name] = SurfacePointCloud(points=pts, normals=nrm)
        return out

    # -- A2: visual frames ----------------------------------------------------------------------
    def setup_fk_functions(self):
        visual_tf = {}
        for link in self.urdf.links:
            if self.collision_link_names is None or link.name in self.collision_link_names:
                xyz, rpy = self.get_link_visual_origin(link)
                vis = sp.rt2tr(sp.rpy2r(np.asarray(rpy).reshape(-1)), np.asarray(xyz).reshape(-1))
                visual_tf[link.name] = _VisualTf(self, link.name, vis)
        return visual_tf

    def compute_fk_surface_points(self, q_user_input, tf_base=None):
        pts_all, nrm_all = [], []
        for name, pc in self.surface_pc_map.items():
            tf = self.visual_tf[name](q_user_input).toarray()
            if tf_base is not None:
                tf = tf_base @ tf
            pts_all.append(pc.points @ tf[:3, :3].T + tf[:3, 3])
            nrm_all.append(pc.normals @ tf[:3, :3].T)
        if not pts_all:
            return np.zeros((0, 3)), np.zeros((0, 3))
        return np.concatenate(pts_all, axis=0), np.concatenate(nrm_all, axis=0)

    def compute_fk_link_surface_points(self, q_user_input, name, tf_base=None):
        tf = self.visual_tf[name](q_user_input).toarray()
        if tf_base is not None:
            tf = tf_base @ tf
        return self.surface_pc_map[name].points @ tf[:3, :3].T + tf[:3, 3]

    # -- A5: voxel field geometry -----------------------------------------------------------------
    def _setup_field(self, lo, hi):
        m, r = self.field_margin, self.grid_resolution
        self.origin = np.array([lo[0] - m, lo[1] - m, lo[2] - m]).reshape((1, 3))
        axes = [np.arange(lo[a] - m, hi[a] + m, r) for a in range(3)]
        grid = np.array(np.meshgrid(*axes, indexing="ij"))
        self.field_shape = grid.shape[1:]
        self.workspace_points = grid.reshape((3, -1)).T
        self.field_size = self.workspace_points.shape[0]
        print("origin", self.origin)
        print("workspace field shape", self.field_shape)
        print("workspace field", self.field_size)

    def setup_workspace_field(self, arm_len, arm_height):
        self.xlim = [0, arm_len]
        self.ylim = [-arm_len, arm_len]
        self.zlim = [0, arm_height + arm_len]
        self._setup_field([self.xlim[0], self.ylim[0], self.zlim[0]], [self.xlim[1], self.ylim[1], self.zlim[1]])

    def setup_points_field(self, points):
        self.workspace_bounds = np.stack((points.min(0), points.max(0)), axis=1)
        self._setup_field(self.workspace_bounds[:, 0], self.workspace_bounds[:, 1])

    def points_to_offsets_numpy(self, points):
        """Clip-then-truncate nearest-node offsets (reference :190-201)."""
        idx = (np.asarray(points, dtype=np.float64) - self.origin) / self.grid_resolution
        for a in range(3):
            idx[:, a] = np.clip(idx[:, a], 0, self.field_shape[a] - 1).astype(np.int32)
        off = idx[:, 2] + self.field_shape[2] * (idx[:, 1] + self.field_shape[1] * idx[:, 0])
        return np.clip(off, 0, self.field_size - 1).astype(np.int32)

    def points_to_offsets(self, points):
        """Floor-then-clamp offsets -- numeric form of the symbolic reference method (:174-187)."""
        idx = np.floor((np.asarray(_as2d(points), dtype=np.float64) - self.origin) / self.grid_resolution)
        for a in range(3):
            idx[:, a] = np.clip(idx[:, a], 0, self.field_shape[a] - 1)
        return (idx[:, 2] + self.field_shape[2] * (idx[:, 1] + self.field_shape[1] * idx[:, 0])).astype(np.int64)

    def compute_plan_cost(self, plan, sdf_cost_obstacle, base_position):
        """Seed-ranking cost (reference :204-215): ``plan`` is ndof-by-T."""
        plan = np.asarray(plan)
        T = plan.shape[1]
        cost = 0
        sdf = np.asarray(sdf_cost_obstacle).reshape(-1)
        for i in range(T):
            pts, _ = self.compute_fk_surface_points(plan[:, i])
            cost += np.sum(sdf[self.points_to_offsets_numpy(pts + np.array(base_position).reshape(1, 3))])
        return cost, np.linalg.norm(plan[:, 0] - plan[:, T - 1])

    # -- occupancy grid for base placement (reference :219-292), numeric ---------------------------------
    def setup_occupancy_grid(self, points, epsilon=0.02):
        xys = points[points[:, 2] > 0.01, :2]
        m, r = self.field_margin, self.grid_resolution
        self.xlim_2d = [0, np.max(xys[:, 0])]
        self.ylim_2d = [np.min(xys[:, 1]), np.max(xys[:, 1])]
        self.occupancy_grid_origin = np.array([self.xlim_2d[0] - m, self.ylim_2d[0] - m]).reshape((1, 2))
        self.xgrid = np.arange(self.xlim_2d[0] - m, self.xlim_2d[1] + m, r)
        self.ygrid = np.arange(self.ylim_2d[0] - m, self.ylim_2d[1] + m, r)
        grid = np.array(np.meshgrid(self.xgrid, self.ygrid, indexing="ij"))
        self.occupancy_grid_shape = grid.shape[1:]
        wp = grid.reshape((2, -1)).T
        dist, _ = KDTree(xys).query(wp)
        self.occupancy_grid_size = wp.shape[0]
        self.occupancy_grid = (dist < epsilon).astype(np.float64)

    def points_to_offsets_occupancy_numpy(self, points):
        idx = np.floor((points[:, :2] - self.occupancy_grid_origin) / self.grid_resolution)
        for a in range(2):
            idx[:, a] = np.clip(idx[:, a], 0, self.occupancy_grid_shape[a] - 1)
        return (idx[:, 1] + self.occupancy_grid_shape[1] * idx[:, 0]).astype(np.int32)

    points_to_offsets_occupancy = points_to_offsets_occupancy_numpy

    def setup_occupancy_grid_function(self):
        def cost(qc, tf_base_inv, occupancy_grid):
            pts, _ = self.compute_fk_surface_points(qc, tf_base=np.asarray(_as2d(tf_base_inv)))
            grid = np.asarray(_as2d(occupancy_grid)).reshape(-1)
            return DM(np.sum(grid[self.points_to_offsets_occupancy_numpy(pts)]))

        return cost

    # -- flattening for the kernels -----------------------------------------------------------------------
    def to_table(self, link_ee: Optional[str] = None, link_gripper: Optional[str] = None) -> RobotTable:
        key = (link_ee, link_gripper)
        if key not in self._tables:
            pts = {name: pc.points for name, pc in self.surface_pc_map.items()}
            self._tables[key] = build_robot_table(self.urdf, self.param_joints, pts, link_ee, link_gripper, name=self.name)
        return self._tables[key]
