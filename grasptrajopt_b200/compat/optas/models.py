"""``optas.RobotModel`` API surface without CasADi (reference: ``optas/models.py:233-1268``).

Numeric (NumPy float64) kinematics only: the symbolic graph the reference builds with CasADi
SX/MX is replaced by the flattened robot table + CUDA kernels (``grasptrajopt_b200.robot_table``,
``csrc/gto_b200.cu``).  Method names, argument order and return conventions follow the
reference so that ``gto`` and ``examples/pybullet_api.py`` run unchanged; results that the
reference returns as ``casadi.DM`` are returned as the ndarray-backed ``DM`` stand-in
(``.toarray()`` works the same).
"""
from __future__ import annotations

import os
import pathlib
from typing import Callable, List, Optional, Sequence

import numpy as np

from grasptrajopt_b200 import spatial as sp
from grasptrajopt_b200.urdf import URDF, Joint, Link, Pose
from .dm import DM, _as2d


class JointTypeNotSupported(NotImplementedError):
    def __init__(self, joint_type: str):
        super().__init__(f"{joint_type} joints are currently not supported")


class Model:
    """Base class (``optas/models.py:79-186``): name, dimension, time derivatives, limits."""

    def __init__(self, name, dim, time_derivs, symbol, dlim, T):
        self.name = name
        self.dim = dim
        self.time_derivs = list(time_derivs)
        self.symbol = symbol
        self.dlim = dlim
        self.T = T

    def get_name(self) -> str:
        return self.name

    def state_name(self, time_deriv: int) -> str:
        assert time_deriv in self.time_derivs, f"Given time derivative {time_deriv=} not recognized, only allowed {self.time_derivs}"
        return self.name + "/" + "d" * time_deriv + self.symbol

    def state_parameter_name(self, time_deriv: int) -> str:
        return self.state_name(time_deriv) + "/p"

    def state_optimized_name(self, time_deriv: int) -> str:
        return self.state_name(time_deriv) + "/x"

    def get_limits(self, time_deriv: int):
        assert time_deriv in self.time_derivs
        assert time_deriv in self.dlim.keys(), f"Limit for time derivative {time_deriv=} has not been given"
        return self.dlim[time_deriv]

    def in_limit(self, x, time_deriv: int):
        lo, up = self.get_limits(time_deriv)
        x = _as2d(x)
        return bool(np.all((_as2d(lo) <= x) & (x <= _as2d(up))))


class TaskModel(Model):
    def __init__(self, name, dim, time_derivs=[0], symbol="y", dlim={}, T=None):
        super().__init__(name, dim, time_derivs, symbol, dlim, T)


class _LinkFunction:
    """Callable returned by the ``get_*_function`` methods: ``f(q)`` for ``n == 1`` and, like the
    reference's ``ListFunction`` / ``Function.map`` (``optas/models.py:749-790``), ``f(Q)`` with
    ``Q`` of shape ndof-by-n otherwise."""

    def __init__(self, fun: Callable, ndof: int, n: int, numpy_output: bool, matrix_out: bool):
        self._fun, self._ndof, self._n = fun, ndof, n
        self._numpy, self._matrix = numpy_output, matrix_out

    def _wrap(self, a: np.ndarray):
        if self._numpy:
            return a.reshape(-1) if a.ndim == 2 and a.shape[1] == 1 else a
        return DM(a)

    def __call__(self, Q):
        Q = _as2d(Q)
        if Q.shape[0] != self._ndof and Q.shape[1] == self._ndof and self._n == 1:
            Q = Q.T
        if self._n == 1:
            return self._wrap(np.asarray(self._fun(Q[:, 0])))
        assert Q.shape[1] == self._n, f"expected input to have shape {self._ndof}-by-{self._n}, got {Q.shape[0]}-by-{Q.shape[1]}"
        outs = [np.asarray(self._fun(Q[:, i])) for i in range(self._n)]
        if self._matrix:
            return [self._wrap(o) for o in outs]
        return self._wrap(np.concatenate([o.reshape(-1, 1) for o in outs], axis=1))


class RobotModel(Model):
    def __init__(
        self,
        urdf_filename: Optional[str] = None,
        urdf_string: Optional[str] = None,
        xacro_filename: Optional[str] = None,
        name: Optional[str] = None,
        time_derivs: List[int] = [0],
        qddlim=None,
        T: Optional[int] = None,
        param_joints: List[str] = [],
    ):
        self.xacro_filename = xacro_filename
        if xacro_filename is not None:
            try:
                import xacro  # optional dependency, as in the reference
            except ImportError as exc:  # pragma: no cover
                raise ImportError("xacro_filename was given but the 'xacro' package is not installed") from exc
            urdf_string = xacro.process(xacro_filename)
        self.urdf = None
        self.urdf_filename = None
        self.urdf_string = None
        if urdf_filename is not None:
            self.urdf_filename = urdf_filename
            self.urdf = URDF.from_xml_file(urdf_filename)
        if urdf_string is not None:
            self.urdf_string = urdf_string
            self.urdf = URDF.from_xml_string(urdf_string)
        assert self.urdf is not None, "You need to supply a urdf, either through filename or as a string"
        self.param_joints = list(param_joints)
        dlim = {
            0: (self.lower_optimized_joint_limits, self.upper_optimized_joint_limits),
            1: (-self.velocity_optimized_joint_limits, self.velocity_optimized_joint_limits),
        }
        if qddlim is not None:
            q = _as2d(qddlim).reshape(-1)
            if q.shape[0] == 1:
                q = q[0] * np.ones(self.ndof)
            assert q.shape[0] == self.ndof, f"expected ddlim to have {self.ndof} elements"
            dlim[2] = (DM(-q), DM(q))
        if name is None:
            name = self.urdf.name
        super().__init__(name, self.ndof, time_derivs, "q", dlim, T)

    # -- URDF access ---------------------------------------------------------------------
    def get_urdf(self):
        return self.urdf

    def get_urdf_dirname(self):
        if self.urdf_filename is not None:
            return pathlib.Path(os.path.dirname(self.urdf_filename))
        if self.xacro_filename is not None:
            return pathlib.Path(os.path.dirname(self.xacro_filename))
        return None

    @property
    def joint_names(self) -> List[str]:
        return [j.name for j in self.urdf.joints]

    @property
    def link_names(self) -> List[str]:
        return [l.name for l in self.urdf.links]

    @property
    def actuated_joint_names(self) -> List[str]:
        return [j.name for j in self.urdf.joints if j.type != "fixed"]

    @property
    def parameter_joint_names(self) -> List[str]:
        return [j for j in self.actuated_joint_names if j in self.param_joints]

    @property
    def optimized_joint_names(self) -> List[str]:
        par = self.parameter_joint_names
        return [j for j in self.actuated_joint_names if j not in par]

    @property
    def optimized_joint_indexes(self) -> List[int]:
        return [self.get_actuated_joint_index(j) for j in self.optimized_joint_names]

    @property
    def parameter_joint_indexes(self) -> List[int]:
        return [self.get_actuated_joint_index(j) for j in self.parameter_joint_names]

    def extract_parameter_dimensions(self, values):
        return DM(_as2d(values)[self.parameter_joint_indexes, :])

    def extract_optimized_dimensions(self, values):
        return DM(_as2d(values)[self.optimized_joint_indexes, :])

    @property
    def ndof(self) -> int:
        return len(self.actuated_joint_names)

    @property
    def num_opt_joints(self) -> int:
        return len(self.optimized_joint_names)

    @property
    def num_param_joints(self) -> int:
        return len(self.parameter_joint_names)

    # -- limits (optas/models.py:438-550) ---------------------------------------------------
    @staticmethod
    def get_joint_lower_limit(joint) -> float:
        return -1e9 if joint.limit is None else joint.limit.lower

    @staticmethod
    def get_joint_upper_limit(joint) -> float:
        return 1e9 if joint.limit is None else joint.limit.upper

    @staticmethod
    def get_velocity_joint_limit(joint) -> float:
        return 1e9 if joint.limit is None else joint.limit.velocity

    def _limits(self, getter, names=None) -> DM:
        vals = [getter(j) for j in self.urdf.joints if (j.type != "fixed" if names is None else j.name in names)]
        return DM(np.array(vals, dtype=np.float64))

    @property
    def lower_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_joint_lower_limit)

    @property
    def upper_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_joint_upper_limit)

    @property
    def velocity_actuated_joint_limits(self) -> DM:
        return self._limits(self.get_velocity_joint_limit)

    @property
    def lower_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_joint_lower_limit, self.optimized_joint_names)

    @property
    def upper_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_joint_upper_limit, self.optimized_joint_names)

    @property
    def velocity_optimized_joint_limits(self) -> DM:
        return self._limits(self.get_velocity_joint_limit, self.optimized_joint_names)

    # -- tree edits / lookups ---------------------------------------------------------------
    def add_base_frame(self, base_link: str, xyz=None, rpy=None, joint_name: str = None) -> None:
        child = self.urdf.get_root()
        xyz = [0.0] * 3 if xyz is None else list(xyz)
        rpy = [0.0] * 3 if rpy is None else list(rpy)
        if not isinstance(joint_name, str):
            joint_name = base_link + "_and_" + child + "_joint"
        self.urdf.add_link(Link(name=base_link))
        self.urdf.add_joint(Joint(name=joint_name, type="fixed", parent=base_link, child=child, origin=Pose(xyz=xyz, rpy=rpy)))

    def get_root_link(self) -> str:
        return self.urdf.get_root()

    def get_link_visual_origin(self, link):
        xyz, rpy = np.zeros(3), np.zeros(3)
        if link.visual is not None and link.visual.origin is not None:
            xyz, rpy = np.array(link.visual.origin.xyz), np.array(link.visual.origin.rpy)
        return DM(xyz), DM(rpy)

    def get_joint_origin(self, joint):
        xyz, rpy = np.zeros(3), np.zeros(3)
        if joint.origin is not None:
            xyz, rpy = np.array(joint.origin.xyz), np.array(joint.origin.rpy)
        return DM(xyz), DM(rpy)

    def get_joint_axis(self, joint) -> DM:
        axis = joint.axis if joint.axis is not None else [1.0, 0.0, 0.0]
        return DM(sp.unit(axis))

    def get_actuated_joint_index(self, joint_name: str) -> int:
        return self.actuated_joint_names.index(joint_name)

    def get_random_joint_positions(self, n: int = 1, xlim=None, ylim=None, zlim=None, base_link=None) -> DM:
        lo = self.lower_actuated_joint_limits.toarray().reshape(-1)
        hi = self.upper_actuated_joint_limits.toarray().reshape(-1)

        def ok(q):
            if not isinstance(base_link, str):
                return True
            for link in self.link_names:
                p = self._link_tf(link, q, base_link)[:3, 3]
                for lim, v in ((xlim, p[0]), (ylim, p[1]), (zlim, p[2])):
                    if lim is not None and not (lim[0] <= v <= lim[1]):
                        return False
            return True

        cols = []
        for _ in range(n):
            q = np.random.uniform(lo, hi)
            while not ok(q):
                q = np.random.uniform(lo, hi)
            cols.append(q.reshape(-1, 1))
        return DM(np.concatenate(cols, axis=1))

    def get_random_pose_in_global_link(self, link_name: str) -> DM:
        return self.get_global_link_transform(link_name, self.get_random_joint_positions())

    # -- numeric kinematics (optas/models.py:826-1268) ------------------------------------------
    def _global_tf(self, link: str, q: np.ndarray) -> np.ndarray:
        assert link in self.urdf.link_map.keys(), f"given link '{link}' does not appear in URDF"
        root = self.urdf.get_root()
        T = np.eye(4)
        if link == root:
            return T
        for joint_name in self.urdf.get_chain(root, link, links=False):
            joint = self.urdf.joint_map[joint_name]
            if joint.origin is not None:
                T = T @ sp.rt2tr(sp.rpy2r(joint.origin.rpy), joint.origin.xyz)
            if joint.type == "fixed":
                continue
            qi = q[self.get_actuated_joint_index(joint.name)]
            axis = sp.unit(joint.axis if joint.axis is not None else [1.0, 0.0, 0.0])
            if joint.type in {"revolute", "continuous"}:
                T = T @ sp.r2t(sp.angvec2r(qi, axis))
            elif joint.type == "prismatic":
                T = T @ sp.rt2tr(np.eye(3), qi * axis)
            else:
                raise JointTypeNotSupported(joint.type)
        return T

    def _link_tf(self, link: str, q: np.ndarray, base_link: str) -> np.ndarray:
        return sp.invt(self._global_tf(base_link, q)) @ self._global_tf(link, q)

    @staticmethod
    def _q(q) -> np.ndarray:
        return _as2d(q).reshape(-1)

    def _per_column(self, fun, q):
        """``listify_output`` semantics (optas/models.py:20-53): a trajectory ndof-by-n gives a list."""
        Q = _as2d(q)
        if Q.shape[1] > 1 and Q.shape[0] == self.ndof:
            return [DM(fun(Q[:, i])) for i in range(Q.shape[1])]
        return DM(fun(Q.reshape(-1)))

    def _function(self, fun, n, numpy_output, matrix_out):
        return _LinkFunction(fun, self.ndof, n, numpy_output, matrix_out)

    def get_global_link_transform(self, link: str, q):
        return self._per_column(lambda v: self._global_tf(link, v), q)

    def get_global_link_transform_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._global_tf(link, v), n, numpy_output, True)

    def get_link_transform(self, link: str, q, base_link: str):
        return self._per_column(lambda v: self._link_tf(link, v, base_link), q)

    def get_link_transform_function(self, link: str, base_link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._link_tf(link, v, base_link), n, numpy_output, True)

    def get_global_link_position(self, link: str, q):
        return self._per_column(lambda v: self._global_tf(link, v)[:3, 3].reshape(3, 1), q)

    def get_global_link_position_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._global_tf(link, v)[:3, 3].reshape(3, 1), n, numpy_output, False)

    def get_link_position(self, link: str, q, base_link: str):
        return self._per_column(lambda v: self._link_tf(link, v, base_link)[:3, 3].reshape(3, 1), q)

    def get_link_position_function(self, link: str, base_link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._link_tf(link, v, base_link)[:3, 3].reshape(3, 1), n, numpy_output, False)

    def get_global_link_rotation(self, link: str, q):
        return self._per_column(lambda v: self._global_tf(link, v)[:3, :3], q)

    def get_global_link_rotation_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._global_tf(link, v)[:3, :3], n, numpy_output, True)

    def get_link_rotation(self, link: str, q, base_link: str):
        return self._per_column(lambda v: self._link_tf(link, v, base_link)[:3, :3], q)

    def get_link_rotation_function(self, link: str, base_link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._link_tf(link, v, base_link)[:3, :3], n, numpy_output, True)

    @staticmethod
    def _quat_xyzw(R: np.ndarray) -> np.ndarray:
        w, x, y, z = sp.mat2quat_wxyz(R)
        return np.array([x, y, z, w]).reshape(4, 1)

    def get_global_link_quaternion(self, link: str, q):
        return self._per_column(lambda v: self._quat_xyzw(self._global_tf(link, v)[:3, :3]), q)

    def get_global_link_quaternion_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._quat_xyzw(self._global_tf(link, v)[:3, :3]), n, numpy_output, False)

    @staticmethod
    def _rpy(R: np.ndarray) -> np.ndarray:
        p = -np.arcsin(np.clip(R[2, 0], -1.0, 1.0))
        r = np.arctan2(R[2, 1], R[2, 2])
        y = np.arctan2(R[1, 0], R[0, 0])
        return np.array([r, p, y]).reshape(3, 1)

    def get_global_link_rpy(self, link: str, q):
        return self._per_column(lambda v: self._rpy(self._global_tf(link, v)[:3, :3]), q)

    def get_global_link_rpy_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._rpy(self._global_tf(link, v)[:3, :3]), n, numpy_output, False)

    def _geometric_jacobian(self, link: str, q: np.ndarray) -> np.ndarray:
        """6-by-ndof geometric Jacobian in the global frame (optas/models.py:1203-1268)."""
        e = self._global_tf(link, q)[:3, 3]
        root = self.urdf.get_root()
        J = np.zeros((6, self.ndof))
        T = np.eye(4)
        for joint_name in self.urdf.get_chain(root, link, links=False):
            joint = self.urdf.joint_map[joint_name]
            if joint.origin is not None:
                T = T @ sp.rt2tr(sp.rpy2r(joint.origin.rpy), joint.origin.xyz)
            if joint.type == "fixed":
                continue
            idx = self.get_actuated_joint_index(joint.name)
            axis = sp.unit(joint.axis if joint.axis is not None else [1.0, 0.0, 0.0])
            z = T[:3, :3] @ axis
            if joint.type in {"revolute", "continuous"}:
                J[:3, idx] = np.cross(z, e - T[:3, 3])
                J[3:, idx] = z
                T = T @ sp.r2t(sp.angvec2r(q[idx], axis))
            elif joint.type == "prismatic":
                J[:3, idx] = z
                T = T @ sp.rt2tr(np.eye(3), q[idx] * axis)
            else:
                raise JointTypeNotSupported(joint.type)
        return J

    def get_global_link_geometric_jacobian(self, link: str, q):
        return self._per_column(lambda v: self._geometric_jacobian(link, v), q)

    def get_global_link_geometric_jacobian_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._geometric_jacobian(link, v), n, numpy_output, True)

    def get_global_link_linear_jacobian(self, link: str, q):
        return self._per_column(lambda v: self._geometric_jacobian(link, v)[:3], q)

    def get_global_link_linear_jacobian_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._geometric_jacobian(link, v)[:3], n, numpy_output, True)

    def get_global_link_angular_geometric_jacobian(self, link: str, q):
        return self._per_column(lambda v: self._geometric_jacobian(link, v)[3:], q)

    def get_global_link_angular_geometric_jacobian_function(self, link: str, n: int = 1, numpy_output: bool = False):
        return self._function(lambda v: self._geometric_jacobian(link, v)[3:], n, numpy_output, True)
