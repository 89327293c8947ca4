"""Flatten a URDF kinematic tree into the table the CUDA kernels consume.

What the reference does symbolically for every (link, knot) pair at NLP-build time
(``optas/models.py:826-868`` chain FK, ``gto/gto_models.py:83-101`` visual frames) is
reduced here, once, to a constant table:

* only *movable* joints that are ancestors of a collision link (or of
  ``link_ee``/``link_gripper``) are kept, in topological order; every run of fixed
  joints is folded into the constant ``origin`` of the next movable joint,
* each collision link carries the constant transform from its last movable joint
  frame to its *visual* frame (link frame x visual origin, ``gto_models.py:96-100``),
* the gripper link additionally carries its plain link frame, because the goal cost
  applies the *link* transform to the gripper point set (``gto/gto_planner.py:79-87``),
* ``G`` = ``link_gripper`` expressed in ``link_ee`` (``gto_planner.py:38``), which is
  constant for every robot the reference ships (only fixed joints in between).

Everything is float64 here; the C-ABI receives float32 copies.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .spatial import rpy2r, rt2tr, angvec2r, invt, unit
from .urdf import URDF

JOINT_REVOLUTE = 1
JOINT_PRISMATIC = 2
MAX_OPT = 16  # optimised joints the kernels support (bitmask width / MMA tile)


@dataclass
class RobotTable:
    name: str
    ndof: int
    actuated_joint_names: List[str]
    opt_qidx: np.ndarray  # [nopt] int32, index into q[ndof]
    par_qidx: np.ndarray  # [npar] int32
    lo: np.ndarray  # [nopt]
    hi: np.ndarray  # [nopt]
    mov_names: List[str]
    mov_parent: np.ndarray  # [nmov] int32
    mov_type: np.ndarray  # [nmov] int32
    mov_origin: np.ndarray  # [nmov,3,4]
    mov_axis: np.ndarray  # [nmov,3]
    mov_qidx: np.ndarray  # [nmov] int32
    mov_opt: np.ndarray  # [nmov] int32 (-1 = parameter joint)
    link_names: List[str]
    link_mov: np.ndarray  # [nlinks] int32
    link_tf: np.ndarray  # [nlinks,3,4]
    link_pt_start: np.ndarray  # [nlinks] int32
    link_pt_count: np.ndarray  # [nlinks] int32
    link_optmask: np.ndarray  # [nlinks] uint32
    points: np.ndarray  # [P,3]
    grip_mov: int = -1
    grip_tf: np.ndarray = field(default_factory=lambda: np.eye(4)[:3])
    grip_pt_start: int = 0
    grip_pt_count: int = 0
    grip_optmask: int = 0
    G: np.ndarray = field(default_factory=lambda: np.eye(4)[:3])
    link_ee: str = ""
    link_gripper: str = ""

    @property
    def nopt(self) -> int:
        return int(self.opt_qidx.shape[0])

    @property
    def nmov(self) -> int:
        return int(self.mov_parent.shape[0])

    @property
    def nlinks(self) -> int:
        return int(self.link_mov.shape[0])

    @property
    def npoints(self) -> int:
        return int(self.points.shape[0])

    # -- persistence (small fixtures that travel to the GPU box) -------------------------
    _ARRAYS = (
        "opt_qidx par_qidx lo hi mov_parent mov_type mov_origin mov_axis mov_qidx mov_opt link_mov link_tf "
        "link_pt_start link_pt_count link_optmask points grip_tf G"
    ).split()

    def save(self, path: str) -> None:
        d = {k: getattr(self, k) for k in self._ARRAYS}
        d["points"] = self.points.astype(np.float32)  # the C-ABI takes float32 points anyway
        d["_meta"] = np.array(
            [
                self.name,
                str(self.ndof),
                ",".join(self.actuated_joint_names),
                ",".join(self.mov_names),
                ",".join(self.link_names),
                str(self.grip_mov),
                str(self.grip_pt_start),
                str(self.grip_pt_count),
                str(self.grip_optmask),
                self.link_ee,
                self.link_gripper,
            ]
        )
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path: str) -> "RobotTable":
        z = np.load(path, allow_pickle=False)
        m = [str(s) for s in z["_meta"]]
        kw = {k: z[k] for k in cls._ARRAYS}
        kw["points"] = kw["points"].astype(np.float64)
        return cls(
            name=m[0],
            ndof=int(m[1]),
            actuated_joint_names=m[2].split(","),
            mov_names=m[3].split(","),
            link_names=m[4].split(","),
            grip_mov=int(m[5]),
            grip_pt_start=int(m[6]),
            grip_pt_count=int(m[7]),
            grip_optmask=int(m[8]),
            link_ee=m[9],
            link_gripper=m[10],
            **kw,
        )


def _joint_origin_tf(joint) -> np.ndarray:
    if joint.origin is None:
        return np.eye(4)
    return rt2tr(rpy2r(joint.origin.rpy), joint.origin.xyz)


def _joint_axis(joint) -> np.ndarray:
    return unit(joint.axis if joint.axis is not None else [1.0, 0.0, 0.0])


def joint_limits(joint):
    if joint.limit is None:
        return -1e9, 1e9
    return float(joint.limit.lower), float(joint.limit.upper)


def build_robot_table(
    urdf: URDF,
    param_joints: Sequence[str],
    point_sets: Dict[str, np.ndarray],
    link_ee: Optional[str] = None,
    link_gripper: Optional[str] = None,
    name: Optional[str] = None,
) -> RobotTable:
    """``point_sets`` maps collision-link name -> [n,3] points in the link's visual frame, in
    the iteration order the planner will use (reference: ``surface_pc_map`` insertion order)."""
    root = urdf.get_root()
    actuated = [j.name for j in urdf.joints if j.type != "fixed"]
    opt_names = [j for j in actuated if j not in set(param_joints)]
    par_names = [j for j in actuated if j in set(param_joints)]
    if len(opt_names) > MAX_OPT:
        raise ValueError(f"{len(opt_names)} optimised joints; the kernels support at most {MAX_OPT}")
    lo, hi = [], []
    for j in urdf.joints:
        if j.name in opt_names:
            l, h = joint_limits(j)
            lo.append(l)
            hi.append(h)

    targets = list(point_sets.keys())
    for extra in (link_ee, link_gripper):
        if extra is not None and extra not in targets:
            targets.append(extra)

    mov_index: Dict[str, int] = {}
    mov_names: List[str] = []
    mov_parent, mov_type, mov_origin, mov_axis, mov_qidx, mov_opt = [], [], [], [], [], []

    def attach(link: str):
        """Returns (movable frame index, constant transform frame->link frame, optmask)."""
        parent_mov = -1
        const = np.eye(4)
        mask = 0
        for jname in urdf.get_chain(root, link, links=False):
            joint = urdf.joint_map[jname]
            const = const @ _joint_origin_tf(joint)
            if joint.type == "fixed":
                continue
            if joint.type in ("revolute", "continuous"):
                jt = JOINT_REVOLUTE
            elif joint.type == "prismatic":
                jt = JOINT_PRISMATIC
            else:
                raise ValueError(f"joint type '{joint.type}' not supported ({jname})")
            if jname not in mov_index:
                mov_index[jname] = len(mov_names)
                mov_names.append(jname)
                mov_parent.append(parent_mov)
                mov_type.append(jt)
                mov_origin.append(const[:3].copy())
                mov_axis.append(_joint_axis(joint))
                mov_qidx.append(actuated.index(jname))
                mov_opt.append(opt_names.index(jname) if jname in opt_names else -1)
            parent_mov = mov_index[jname]
            if mov_opt[parent_mov] >= 0:
                mask |= 1 << mov_opt[parent_mov]
            const = np.eye(4)
        return parent_mov, const, mask

    link_names, link_mov, link_tf, link_start, link_count, link_mask, pts = [], [], [], [], [], [], []
    cursor = 0
    for lname, p in point_sets.items():
        p = np.asarray(p, dtype=np.float64).reshape(-1, 3)
        mov, const, mask = attach(lname)
        link = urdf.link_map[lname]
        vis = np.eye(4)
        if link.visual is not None and link.visual.origin is not None:
            vis = rt2tr(rpy2r(link.visual.origin.rpy), link.visual.origin.xyz)
        link_names.append(lname)
        link_mov.append(mov)
        link_tf.append((const @ vis)[:3])
        link_start.append(cursor)
        link_count.append(p.shape[0])
        link_mask.append(mask)
        pts.append(p)
        cursor += p.shape[0]

    table = RobotTable(
        name=name or urdf.name,
        ndof=len(actuated),
        actuated_joint_names=actuated,
        opt_qidx=np.array([actuated.index(j) for j in opt_names], dtype=np.int32),
        par_qidx=np.array([actuated.index(j) for j in par_names], dtype=np.int32),
        lo=np.array(lo, dtype=np.float64),
        hi=np.array(hi, dtype=np.float64),
        mov_names=mov_names,
        mov_parent=np.zeros(0, np.int32),
        mov_type=np.zeros(0, np.int32),
        mov_origin=np.zeros((0, 3, 4)),
        mov_axis=np.zeros((0, 3)),
        mov_qidx=np.zeros(0, np.int32),
        mov_opt=np.zeros(0, np.int32),
        link_names=link_names,
        link_mov=np.array(link_mov, dtype=np.int32),
        link_tf=np.array(link_tf, dtype=np.float64).reshape(-1, 3, 4),
        link_pt_start=np.array(link_start, dtype=np.int32),
        link_pt_count=np.array(link_count, dtype=np.int32),
        link_optmask=np.array(link_mask, dtype=np.uint32),
        points=np.concatenate(pts, axis=0) if pts else np.zeros((0, 3)),
        link_ee=link_ee or "",
        link_gripper=link_gripper or "",
    )

    if link_gripper is not None:
        if link_gripper not in point_sets:
            raise ValueError(f"link_gripper '{link_gripper}' must be one of the collision links (it supplies the goal point set)")
        mov, const, mask = attach(link_gripper)
        table.grip_mov = mov
        table.grip_tf = const[:3].copy()
        table.grip_optmask = mask
        k = link_names.index(link_gripper)
        table.grip_pt_start = int(link_start[k])
        table.grip_pt_count = int(link_count[k])
        ee = link_ee or link_gripper
        mov_e, const_e, _ = attach(ee)
        if mov_e != mov:
            raise ValueError(
                f"link_ee '{ee}' and link_gripper '{link_gripper}' are separated by a movable joint; "
                "the goal transform G would depend on q (not supported)"
            )
        table.G = (invt(const_e) @ const)[:3].copy()

    table.mov_parent = np.array(mov_parent, dtype=np.int32)
    table.mov_type = np.array(mov_type, dtype=np.int32)
    table.mov_origin = np.array(mov_origin, dtype=np.float64).reshape(-1, 3, 4)
    table.mov_axis = np.array(mov_axis, dtype=np.float64).reshape(-1, 3)
    table.mov_qidx = np.array(mov_qidx, dtype=np.int32)
    table.mov_opt = np.array(mov_opt, dtype=np.int32)
    return table


def prepend_planar_base(table: RobotTable, xlim=(-0.6, 0.6), ylim=(-0.6, 0.6), yawlim=(-np.pi, np.pi)) -> RobotTable:
    """"10-DoF mobile" variant (BASELINE config C4): the base pose (x, y, yaw) becomes three
    virtual optimised joints in front of the tree -- the single-trajectory analogue of the
    reference's ``BasePlanner`` ``TaskModel(dim=3)`` (``gto/base_planner.py:23,44-51``)."""
    nb = 3
    if table.nopt + nb > MAX_OPT:
        raise ValueError("too many optimised joints")
    I34 = np.eye(4)[:3]
    t = RobotTable(**{k: (v.copy() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v)) for k, v in table.__dict__.items()})
    t.name = table.name + "_mobile"
    t.ndof = table.ndof + nb
    t.actuated_joint_names = ["base_x", "base_y", "base_yaw"] + list(table.actuated_joint_names)
    t.opt_qidx = np.concatenate([np.arange(nb), table.opt_qidx + nb]).astype(np.int32)
    t.par_qidx = (table.par_qidx + nb).astype(np.int32)
    t.lo = np.concatenate([[xlim[0], ylim[0], yawlim[0]], table.lo])
    t.hi = np.concatenate([[xlim[1], ylim[1], yawlim[1]], table.hi])
    t.mov_names = ["base_x", "base_y", "base_yaw"] + list(table.mov_names)
    t.mov_parent = np.concatenate([[-1, 0, 1], np.where(table.mov_parent < 0, 2, table.mov_parent + nb)]).astype(np.int32)
    t.mov_type = np.concatenate([[JOINT_PRISMATIC, JOINT_PRISMATIC, JOINT_REVOLUTE], table.mov_type]).astype(np.int32)
    t.mov_origin = np.concatenate([np.stack([I34, I34, I34]), table.mov_origin], axis=0)
    t.mov_axis = np.concatenate([np.eye(3), table.mov_axis], axis=0)
    t.mov_qidx = np.concatenate([np.arange(nb), table.mov_qidx + nb]).astype(np.int32)
    t.mov_opt = np.concatenate([np.arange(nb), np.where(table.mov_opt < 0, -1, table.mov_opt + nb)]).astype(np.int32)
    t.link_mov = np.where(table.link_mov < 0, 2, table.link_mov + nb).astype(np.int32)
    t.link_optmask = ((table.link_optmask.astype(np.uint64) << nb) | 0b111).astype(np.uint32)
    t.grip_mov = 2 if table.grip_mov < 0 else table.grip_mov + nb
    t.grip_optmask = (int(table.grip_optmask) << nb) | 0b111
    return t
