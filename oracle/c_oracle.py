"""ctypes loader of oracle/libgto_oracle.so (the C restatement of the oracle).  TEST INFRASTRUCTURE ONLY: imported by
tests/ and by the cpu_baseline / --impl reference legs of bench.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from grasptrajopt_b200 import capi  # struct definitions of include/gto_b200.h  # noqa: E402

_dp, _fp, _ip, _up = capi._dp, capi._fp, capi._ip, capi._up


class OracleField(C.Structure):
    _fields_ = [("cost", _fp), ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("ox", C.c_double), ("oy", C.c_double),
                ("oz", C.c_double), ("pitch", C.c_double)]


_lib = None


def load(build=True):
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libgto_oracle.so")
        srcs = [os.path.join(HERE, f) for f in ("gto_oracle.c", "base_oracle.c")]
        if build and (not os.path.exists(path) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in srcs)):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(path)
        _lib.oracle_solve_batch.argtypes = [C.POINTER(capi.RobotDesc), C.POINTER(OracleField), C.POINTER(capi.BatchIn), C.POINTER(capi.Options),
                                            C.POINTER(capi.BatchOut), C.c_int]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_base_place.argtypes = [C.POINTER(capi.RobotDesc), C.POINTER(capi.BaseIn), C.POINTER(capi.Options), C.POINTER(capi.BaseOut), C.c_int]
    return _lib


def default_options(**kw):
    """Same defaults as gto_default_options / gto_oracle.SolverOptions (no libgto_b200 needed)."""
    o = capi.Options()
    vals = dict(max_iter=100, tol_step=1e-6, tol_grad=1e-6, lambda0=1e-3, lambda_min=1e-9, lambda_max=1e9, eta=1e-4, noise_rel=1e-6,
                bound_eps=1e-12, check_every=4, ftol=1e-6, lambda_slow=1e30, slow_window=0, slow_ftol=1e-3, as_rounds=1, lambda_reject=1e-4, lambda_conv=1e-2, bundle=3, bundle_radius=3e-3)
    vals.update(kw)
    for k, v in vals.items():
        setattr(o, k, v)
    return o


def _robot_desc(t):
    a = dict(
        opt_qidx=np.ascontiguousarray(t.opt_qidx, np.int32), lo=np.ascontiguousarray(t.lo, np.float64), hi=np.ascontiguousarray(t.hi, np.float64),
        mov_parent=np.ascontiguousarray(t.mov_parent, np.int32), mov_type=np.ascontiguousarray(t.mov_type, np.int32),
        mov_origin=np.ascontiguousarray(t.mov_origin, np.float64), mov_axis=np.ascontiguousarray(t.mov_axis, np.float64),
        mov_qidx=np.ascontiguousarray(t.mov_qidx, np.int32), mov_opt=np.ascontiguousarray(t.mov_opt, np.int32),
        link_mov=np.ascontiguousarray(t.link_mov, np.int32), link_tf=np.ascontiguousarray(t.link_tf, np.float64),
        link_pt_start=np.ascontiguousarray(t.link_pt_start, np.int32), link_pt_count=np.ascontiguousarray(t.link_pt_count, np.int32),
        link_optmask=np.ascontiguousarray(t.link_optmask, np.uint32), points=np.ascontiguousarray(t.points, np.float32),
        grip_tf=np.ascontiguousarray(t.grip_tf, np.float64))
    d = capi.RobotDesc()
    d.ndof, d.nopt, d.nmov, d.nlinks, d.npoints = t.ndof, t.nopt, t.nmov, t.nlinks, t.npoints
    for k in ("opt_qidx", "mov_parent", "mov_type", "mov_qidx", "mov_opt", "link_mov", "link_pt_start", "link_pt_count"):
        setattr(d, k, a[k].ctypes.data_as(_ip))
    for k in ("lo", "hi", "mov_origin", "mov_axis", "link_tf", "grip_tf"):
        setattr(d, k, a[k].ctypes.data_as(_dp))
    d.link_optmask = a["link_optmask"].ctypes.data_as(_up)
    d.points = a["points"].ctypes.data_as(_fp)
    d.grip_mov, d.grip_pt_start, d.grip_pt_count, d.grip_optmask = int(t.grip_mov), int(t.grip_pt_start), int(t.grip_pt_count), int(t.grip_optmask)
    return d, a


def solve_workload(w, indices=None, nthreads=0, options=None):
    """Solve problems `indices` of a Workload with the C oracle; returns dict like GtoContext.solve_batch."""
    from grasptrajopt_b200.workloads import slice_batch

    lib = load()
    t = w.table
    b = w.batch
    if indices is not None:
        idx = np.asarray(indices)
        b = capi.Batch(T=b.T, dt=b.dt, qc=b.qc[idx], q_seed=b.q_seed[idx], goal_tf=b.goal_tf[idx],
                       base_position=None if b.base_position is None else b.base_position[idx],
                       field_all=None if b.field_all is None else b.field_all[idx], field_obs=None if b.field_obs is None else b.field_obs[idx],
                       standoff_offset=b.standoff_offset, use_standoff=b.use_standoff, collision_avoidance=b.collision_avoidance,
                       w_goal=b.w_goal, w_obs=b.w_obs, w_vel=b.w_vel)
    B, T = b.B, int(b.T)
    d, keep_r = _robot_desc(t)
    nslots = (max(w.fields) + 1) if w.fields else 1
    farr = (OracleField * nslots)()
    keep_f = []
    for s, cf in w.fields.items():
        c = np.ascontiguousarray(cf.cost, np.float32)
        keep_f.append(c)
        farr[s].cost = c.ctypes.data_as(_fp)
        farr[s].nx, farr[s].ny, farr[s].nz = c.shape
        farr[s].ox, farr[s].oy, farr[s].oz = (float(v) for v in cf.origin)
        farr[s].pitch = float(cf.pitch)
    keep = dict(qc=np.ascontiguousarray(b.qc, np.float64), q_seed=np.ascontiguousarray(b.q_seed, np.float64),
                goal_tf=np.ascontiguousarray(b.goal_tf, np.float64).reshape(B, 24),
                base=np.ascontiguousarray(b.base_position if b.base_position is not None else np.zeros((B, 3)), np.float64),
                fa=np.ascontiguousarray(b.field_all if b.field_all is not None else -np.ones(B), np.int32),
                fo=np.ascontiguousarray(b.field_obs if b.field_obs is not None else -np.ones(B), np.int32))
    s = capi.BatchIn()
    s.B, s.T, s.dt = B, T, float(b.dt)
    s.qc, s.q_seed, s.goal_tf, s.base_position = (keep[k].ctypes.data_as(_dp) for k in ("qc", "q_seed", "goal_tf", "base"))
    s.field_all, s.field_obs = keep["fa"].ctypes.data_as(_ip), keep["fo"].ctypes.data_as(_ip)
    s.standoff_offset, s.use_standoff, s.collision_avoidance = int(b.standoff_offset), int(bool(b.use_standoff)), int(bool(b.collision_avoidance))
    s.w_goal, s.w_obs, s.w_vel, s.flags = float(b.w_goal), float(b.w_obs), float(b.w_vel), int(getattr(w.batch, 'flags', 0)) & capi.FLAG_OBS_LINEAR
    res = dict(Q=np.zeros((B, T, t.ndof)), dQ=np.zeros((B, T - 1, t.ndof)), cost=np.zeros(B), iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32))
    o = capi.BatchOut()
    o.Q, o.dQ, o.cost = res["Q"].ctypes.data_as(_dp), res["dQ"].ctypes.data_as(_dp), res["cost"].ctypes.data_as(_dp)
    o.iters, o.status = res["iters"].ctypes.data_as(_ip), res["status"].ctypes.data_as(_ip)
    opts = options if options is not None else default_options()
    rc = lib.oracle_solve_batch(C.byref(d), farr, C.byref(s), C.byref(opts), C.byref(o), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_solve_batch failed ({rc})")
    res["threads"] = lib.oracle_num_threads() if nthreads <= 0 else nthreads
    return res


def base_place(t, qc, RTs, w_effort=0.01, occupancy=None, occ_origin=(0.0, 0.0), occ_resolution=0.05, nthreads=0, options=None):
    """C restatement of oracle/base_oracle.py (oracle/base_oracle.c), same result layout as GtoContext.base_place."""
    lib = load()
    RTs = np.asarray(RTs, dtype=np.float64)
    if RTs.ndim == 3:
        RTs = RTs[None]
    B, n = RTs.shape[:2]
    G = np.eye(4)
    G[:3] = t.G
    goal = np.ascontiguousarray((RTs @ G)[:, :, :3, :], np.float64)
    qcv = np.ascontiguousarray(np.asarray(qc, np.float64).reshape(-1))
    d, keep_r = _robot_desc(t)
    bi = capi.BaseIn()
    bi.B, bi.n_goals, bi.qc, bi.goal_tf, bi.w_effort = B, n, qcv.ctypes.data_as(_dp), goal.ctypes.data_as(_dp), float(w_effort)
    occ = None
    if occupancy is not None:
        occ = np.ascontiguousarray(occupancy, np.float32)
        bi.occupancy = occ.ctypes.data_as(_fp)
        bi.occ_dims[0], bi.occ_dims[1] = occ.shape
        bi.occ_origin[0], bi.occ_origin[1] = float(occ_origin[0]), float(occ_origin[1])
        bi.occ_resolution = float(occ_resolution)
    res = dict(Q=np.zeros((B, n, t.ndof)), y=np.zeros((B, 3)), cost=np.zeros(B), collision=np.zeros(B), iters=np.zeros(B, np.int32),
               status=np.zeros(B, np.int32))
    bo = capi.BaseOut()
    bo.Q, bo.y, bo.cost, bo.collision = (res[k].ctypes.data_as(_dp) for k in ("Q", "y", "cost", "collision"))
    bo.iters, bo.status = res["iters"].ctypes.data_as(_ip), res["status"].ctypes.data_as(_ip)
    opts = options if options is not None else default_options()
    rc = lib.oracle_base_place(C.byref(d), C.byref(bi), C.byref(opts), C.byref(bo), int(nthreads))
    if rc < 1:
        raise RuntimeError(f"oracle_base_place failed ({rc})")
    res["threads"] = rc
    return res
