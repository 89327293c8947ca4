"""The BASELINE.json configurations as concrete, seeded inputs.

  C1  Panda, single tabletop problem, T=30               (the reference's CPU-runnable case)
  C2  Panda, 256 candidate grasps x 30 knots, P=2000, 128^3 field, 1 GPU     <- headline metric
  C3  Fetch-8, 1024 grasps x 50 knots, P=4000, 256^3 shelf field
  C4  Fetch-10 (mobile), 4096 seeds x 50 knots, P=4000, 128^3, sharded over 8 GPUs
  C5  Panda clutter, 16384 problems x 30 knots, 64 scenes x 256, 128^3, scaling sweep

Everything is generated from ``np.random.default_rng`` with a fixed seed, so the CUDA path, the oracle and
the CPU baseline see identical bits.  ``scale`` < 1 shrinks the batch (and nothing else) for tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import scenes as S
from .capi import Batch, goal_transforms
from .robot_table import RobotTable

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

PANDA_QC = np.array([0.0, -1.285, 0.0, -2.356, 0.0, 1.571, 0.785, 0.0, 0.0])  # data/configs/panda.yaml:15
FETCH_QC = np.array([0.0, 0.0, 0.38, 0.009195, 0.908270, 1.32, 0.7, 0.0, -2.0, 0.0, -0.57, 0.0, 0.0, 0.0, 0.0])  # fetch.yaml:15


@dataclass
class Workload:
    name: str
    table: RobotTable
    fields: Dict[int, S.CostField]  # slot -> field
    batch: Batch
    RT: np.ndarray  # [B,4,4] goal poses (robot base frame)
    standoff_distance: float
    axis_standoff: str
    q_star: np.ndarray  # [B,ndof] configurations the goals were generated from
    description: str = ""
    field_specs: Dict[int, tuple] = field(default_factory=dict)  # slot -> (boxes, lo, hi, n): what the field was built from


def _table(name: str) -> RobotTable:
    return RobotTable.load(os.path.join(ASSETS, name + ".npz"))


def _assemble(name, table, fields, qc, RT, QG, T, field_all, field_obs, standoff_distance, axis, description, q_star,
              qc_noise=0.0, rng=None, Tmax=10.0) -> Workload:
    B = RT.shape[0]
    dt = Tmax / (T - 1)  # gto/gto_planner.py:25-28 (Tmax = 10 s over T knots)
    QC = np.tile(qc, (B, 1))
    if qc_noise > 0:
        QC[:, table.opt_qidx] = np.clip(QC[:, table.opt_qidx] + rng.normal(0, qc_noise, size=(B, table.nopt)), table.lo, table.hi)
    seeds = np.stack([S.smoothstep_seed(QC[b], QG[b : b + 1], T)[0] for b in range(B)])
    seeds[:, :, table.par_qidx] = QC[:, None, table.par_qidx]  # parameter rows <- qc (gto_planner.py:155-158)
    batch = Batch(T=T, dt=dt, qc=QC, q_seed=seeds, goal_tf=goal_transforms(table, RT, standoff_distance, axis),
                  base_position=np.zeros((B, 3)), field_all=np.asarray(field_all, np.int32), field_obs=np.asarray(field_obs, np.int32),
                  standoff_offset=-10, use_standoff=True, collision_avoidance=True)
    return Workload(name, table, fields, batch, RT, standoff_distance, axis, q_star, description)


def _obj_box(center, size=(0.06, 0.06, 0.12)):
    return [((float(center[0]), float(center[1]), float(center[2])), tuple(size))]


def make_workload(config: str, scale: float = 1.0, seed: Optional[int] = None, n_field: Optional[int] = None, cache: bool = True) -> Workload:
    """The seeded workload of a BASELINE configuration (see ``_build_workload``).  Rejection sampling of thousands of reachable grasps
    takes tens of seconds on the host (C4: 38 s, C5: 51 s), so large workloads are kept in the temp directory -- everything but the
    cost fields, which are rebuilt from their boxes (milliseconds) -- and shared by later runs and by the ranks of one run."""
    import pickle
    import tempfile

    full = {"C1": 1, "C2": 256, "C3": 1024, "C4": 4096, "C5": 16384}[config.upper()]
    if not cache or full * scale < 1000:
        return _build_workload(config, scale, seed, n_field)
    path = os.path.join(tempfile.gettempdir(), f"gto_b200_workload_v3_{config.upper()}_{scale:.6f}_{seed}_{n_field}.pkl")
    if os.path.exists(path):
        try:
            with open(path, "rb") as fh:
                w = pickle.load(fh)
            w.fields = {slot: S.make_field(*spec) for slot, spec in w.field_specs.items()}
            return w
        except Exception:
            pass
    w = _build_workload(config, scale, seed, n_field)
    try:
        fields, w.fields = w.fields, {}
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "wb") as fh:
            pickle.dump(w, fh, protocol=4)
        os.replace(tmp, path)
    except OSError:
        pass
    finally:
        w.fields = fields
    return w


def _build_workload(config: str, scale: float = 1.0, seed: Optional[int] = None, n_field: Optional[int] = None) -> Workload:
    """Scenes follow the reference pipeline (``examples/pybullet_gto_planning.py:178-190``): ``sdf_cost_all`` (slot 2s)
    contains every obstacle *and* the target object, ``sdf_cost_obstacle`` (slot 2s+1) everything but the target; the
    first T-10 knots read the former, the approach knots the latter (``gto/gto_planner.py:117-131``)."""
    config = config.upper()
    idx = {"C1": 1, "C2": 2, "C3": 3, "C4": 4, "C5": 5}[config]
    rng = np.random.default_rng(idx if seed is None else seed)
    specs = {}

    def done(w, slot_boxes, lo_, hi_, n_):
        w.field_specs = {slot: (list(bx), np.asarray(lo_), np.asarray(hi_), n_) for slot, bx in slot_boxes.items()}
        return w

    if config in ("C1", "C2"):
        t = _table("panda_c2")
        B = 1 if config == "C1" else max(1, int(round(256 * scale)))
        n = n_field or 128
        target = np.array([0.5, 0.0, 0.06])
        boxes = S.tabletop_boxes() + [((0.5, 0.25, 0.075), (0.08, 0.08, 0.15)), ((0.35, -0.22, 0.05), (0.1, 0.1, 0.1)),
                                      ((0.68, -0.1, 0.1), (0.06, 0.06, 0.2))]
        lo, hi = S.workspace_box(1.0, 0.0)
        f_obs = S.make_field(boxes, lo, hi, n)
        f_all = S.make_field(boxes + _obj_box(target), lo, hi, n)
        specs = {0: boxes + _obj_box(target), 1: boxes}
        RT, QS, QG = S.sample_grasps_around(t, PANDA_QC, target, B, rng, approach_axis="z", reach=0.10, boxes=boxes)
        return done(_assemble(config, t, {0: f_all, 1: f_obs}, PANDA_QC, RT, QG, 30, [0] * B, [1] * B, -0.1, "z",
                         f"Panda tabletop, {B} candidate grasps x 30 knots, P={t.npoints}, {n}^3 SDF", QS, rng=rng), specs, lo, hi, n)
    if config == "C3":
        t = _table("fetch8_c3")
        B = max(1, int(round(1024 * scale)))
        n = n_field or 256
        x0, z0 = 0.95, 0.55
        boxes = S.shelf_boxes(x0=x0, z0=z0)
        target = np.array([x0 - 0.11, 0.0, z0 + 0.3905 + 0.011 + 0.06])
        lo, hi = S.workspace_box(1.1, 1.1)
        f_obs = S.make_field(boxes, lo, hi, n)
        f_all = S.make_field(boxes + _obj_box(target), lo, hi, n)
        specs = {0: boxes + _obj_box(target), 1: boxes}
        RT, QS, QG = S.sample_grasps_around(t, FETCH_QC, target, B, rng, approach_axis="x", reach=0.0, approach_dir=(1, 0, 0),
                                            min_cos=0.6, boxes=boxes)
        return done(_assemble(config, t, {0: f_all, 1: f_obs}, FETCH_QC, RT, QG, 50, [0] * B, [1] * B, -0.2, "x",
                         f"Fetch 8-DoF shelf, {B} grasps x 50 knots, P={t.npoints}, {n}^3 SDF", QS, rng=rng), specs, lo, hi, n)
    if config == "C4":
        t = _table("fetch10_c4")
        B = max(1, int(round(4096 * scale)))
        n = n_field or 128
        boxes = S.tabletop_boxes(z_top=0.75, x=(0.7, 1.5), y=(-0.6, 0.6))
        target = np.array([0.9, 0.0, 0.81])
        lo, hi = np.array([-0.75, -1.5, 0.0]), np.array([1.6, 1.5, 2.2])
        f_obs = S.make_field(boxes, lo, hi, n)
        f_all = S.make_field(boxes + _obj_box(target), lo, hi, n)
        specs = {0: boxes + _obj_box(target), 1: boxes}
        qc = np.concatenate([[0.0, 0.0, 0.0], FETCH_QC])
        RT, QS, QG = S.sample_grasps_around(t, qc, target, B, rng, approach_axis="x", reach=0.0, approach_dir=(0, 0, -1), min_cos=0.2,
                                            boxes=boxes)
        return done(_assemble(config, t, {0: f_all, 1: f_obs}, qc, RT, QG, 50, [0] * B, [1] * B, -0.1, "x",
                         f"Fetch mobile 10-DoF tabletop, {B} seeds x 50 knots, P={t.npoints}, {n}^3 SDF", QS, qc_noise=0.1, rng=rng), specs, lo, hi, n)
    if config == "C5":
        t = _table("panda_c2")
        nscene = max(1, int(round(64 * scale)))
        per = 256
        n = n_field or 128
        lo, hi = S.workspace_box(1.0, 0.0)
        fields, RTs, QSs, QGs, fa, fo = {}, [], [], [], [], []
        for s in range(nscene):
            boxes = S.clutter_boxes(rng)
            target = np.array([rng.uniform(0.38, 0.65), rng.uniform(-0.3, 0.3), 0.26])
            boxes_t = boxes + [((target[0], target[1], 0.10), (0.08, 0.08, 0.20))]  # pedestal keeps the target above the clutter
            fields[2 * s] = S.make_field(boxes_t + _obj_box(target), lo, hi, n)
            fields[2 * s + 1] = S.make_field(boxes_t, lo, hi, n)
            specs[2 * s], specs[2 * s + 1] = boxes_t + _obj_box(target), boxes_t
            RT, QS, QG = S.sample_grasps_around(t, PANDA_QC, target, per, rng, approach_axis="z", reach=0.10, boxes=boxes_t)
            RTs.append(RT); QSs.append(QS); QGs.append(QG)
            fa += [2 * s] * per
            fo += [2 * s + 1] * per
        RT, QS, QG = np.concatenate(RTs), np.concatenate(QSs), np.concatenate(QGs)
        return done(_assemble(config, t, fields, PANDA_QC, RT, QG, 30, fa, fo, -0.1, "z",
                         f"Panda clutter, {nscene} scenes x {per} problems x 30 knots, P={t.npoints}, {n}^3 SDF", QS, qc_noise=0.1, rng=rng), specs, lo, hi, n)
    raise ValueError(config)


def slice_batch(b: Batch, lo: int, hi: int) -> Batch:
    """Problems [lo, hi) of a batch (static contiguous shard)."""
    return Batch(T=b.T, dt=b.dt, qc=b.qc[lo:hi], q_seed=b.q_seed[lo:hi], goal_tf=b.goal_tf[lo:hi],
                 base_position=None if b.base_position is None else b.base_position[lo:hi],
                 field_all=None if b.field_all is None else b.field_all[lo:hi], field_obs=None if b.field_obs is None else b.field_obs[lo:hi],
                 standoff_offset=b.standoff_offset, use_standoff=b.use_standoff, collision_avoidance=b.collision_avoidance,
                 w_goal=b.w_goal, w_obs=b.w_obs, w_vel=b.w_vel, flags=b.flags)


def shard_range(B: int, rank: int, world: int):
    """Contiguous equal split of [0,B) (remainder to the first ranks)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
