"""ctypes binding of ``libgto_b200.so`` (``include/gto_b200.h``).

This is the stub a maintainer of the reference would add to call the CUDA path from Python instead of
``optas.CasADiSolver`` -> ``casadi.nlpsol`` (``optas/solver.py:384-400``); see INTEGRATION.md.  There is no CPU
fallback: a missing library raises ``GtoLibraryError`` and a missing GPU raises ``GtoError`` from ``gto_create``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .robot_table import RobotTable

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgto_b200.so")

GTO_OK = 0
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_NAN, STATUS_STALLED, STATUS_SLOW = 0, 1, 2, 3, 4
FLAG_NO_JROWS, FLAG_OBS_LINEAR, FLAG_NO_CULL = 1, 2, 32

SYMBOLS = [
    "gto_abi_version", "gto_create", "gto_destroy", "gto_last_error", "gto_default_options", "gto_set_robot",
    "gto_set_field", "gto_solve_batch", "gto_upload_batch", "gto_solve_resident", "gto_download_batch",
    "gto_result_device_ptr", "gto_eval_batch", "gto_get_profile", "gto_configure", "gto_plan_cost", "gto_cloud_set", "gto_cloud_query",
    "gto_cloud_backproject", "gto_base_place",
]


class GtoLibraryError(RuntimeError):
    pass


class GtoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgto_b200 error {code}: {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)


class RobotDesc(C.Structure):
    _fields_ = [
        ("ndof", C.c_int32), ("nopt", C.c_int32), ("opt_qidx", _ip), ("lo", _dp), ("hi", _dp),
        ("nmov", C.c_int32), ("mov_parent", _ip), ("mov_type", _ip), ("mov_origin", _dp), ("mov_axis", _dp),
        ("mov_qidx", _ip), ("mov_opt", _ip),
        ("nlinks", C.c_int32), ("link_mov", _ip), ("link_tf", _dp), ("link_pt_start", _ip), ("link_pt_count", _ip),
        ("link_optmask", _up),
        ("npoints", C.c_int32), ("points", _fp),
        ("grip_mov", C.c_int32), ("grip_tf", _dp), ("grip_pt_start", C.c_int32), ("grip_pt_count", C.c_int32),
        ("grip_optmask", C.c_uint32),
    ]


class Options(C.Structure):
    _fields_ = [
        ("max_iter", C.c_int32), ("tol_step", C.c_double), ("tol_grad", C.c_double), ("lambda0", C.c_double),
        ("lambda_min", C.c_double), ("lambda_max", C.c_double), ("eta", C.c_double), ("noise_rel", C.c_double),
        ("bound_eps", C.c_double), ("check_every", C.c_int32), ("ftol", C.c_double), ("lambda_slow", C.c_double),
        ("slow_window", C.c_int32), ("slow_ftol", C.c_double), ("as_rounds", C.c_int32), ("lambda_reject", C.c_double), ("lambda_conv", C.c_double),
        ("bundle", C.c_int32), ("bundle_radius", C.c_double),
    ]


class BatchIn(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("T", C.c_int32), ("dt", C.c_double), ("qc", _dp), ("q_seed", _dp), ("goal_tf", _dp),
        ("base_position", _dp), ("field_all", _ip), ("field_obs", _ip), ("standoff_offset", C.c_int32),
        ("use_standoff", C.c_int32), ("collision_avoidance", C.c_int32), ("w_goal", C.c_double), ("w_obs", C.c_double),
        ("w_vel", C.c_double), ("flags", C.c_uint32),
    ]


class BatchOut(C.Structure):
    _fields_ = [("Q", _dp), ("dQ", _dp), ("cost", _dp), ("iters", _ip), ("status", _ip)]


class EvalOut(C.Structure):
    _fields_ = [("rows", _fp), ("H", _fp), ("g", _dp), ("cost", _dp)]


class BaseIn(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("n_goals", C.c_int32), ("qc", _dp), ("goal_tf", _dp), ("w_effort", C.c_double), ("occupancy", _fp),
        ("occ_dims", C.c_int32 * 2), ("occ_origin", C.c_double * 2), ("occ_resolution", C.c_double),
    ]


class BaseOut(C.Structure):
    _fields_ = [("Q", _dp), ("y", _dp), ("cost", _dp), ("collision", _dp), ("iters", _ip), ("status", _ip)]


class Profile(C.Structure):
    _fields_ = [
        ("solve_ms", C.c_double), ("linearize_ms", C.c_double), ("step_ms", C.c_double),
        ("linearize_launches", C.c_int32), ("step_launches", C.c_int32), ("iterations", C.c_int32),
        ("knot_items", C.c_int64), ("jrow_bytes", C.c_int64),
        ("problem_iterations", C.c_int64), ("linearize_launches_with_work", C.c_int32), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("links_tested", C.c_int64), ("links_active", C.c_int64), ("kernel_launches", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load_library(path: Optional[str] = None):
    """Load the shared library (never falls back to anything else)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("GTO_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise GtoLibraryError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
    lib = C.CDLL(path)
    lib.gto_abi_version.restype = C.c_int
    lib.gto_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.gto_destroy.argtypes = [C.c_void_p]
    lib.gto_destroy.restype = None
    lib.gto_last_error.argtypes = [C.c_void_p]
    lib.gto_last_error.restype = C.c_char_p
    lib.gto_default_options.argtypes = [C.POINTER(Options)]
    lib.gto_default_options.restype = None
    lib.gto_set_robot.argtypes = [C.c_void_p, C.POINTER(RobotDesc)]
    lib.gto_set_field.argtypes = [C.c_void_p, C.c_int, _fp, _ip, _dp, C.c_double]
    lib.gto_solve_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(Options), C.POINTER(BatchOut)]
    lib.gto_upload_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn)]
    lib.gto_solve_resident.argtypes = [C.c_void_p, C.POINTER(Options)]
    lib.gto_download_batch.argtypes = [C.c_void_p, C.POINTER(BatchOut)]
    lib.gto_result_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.gto_eval_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(EvalOut)]
    lib.gto_get_profile.argtypes = [C.c_void_p, C.POINTER(Profile)]
    lib.gto_configure.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.gto_cloud_set.argtypes = [C.c_void_p, _dp, C.c_int64]
    lib.gto_cloud_query.argtypes = [C.c_void_p, _dp, C.c_int64, _fp, C.c_int32, C.c_int32, _dp, _dp, C.c_int32, C.c_double, C.c_double, _fp, _dp]
    lib.gto_cloud_backproject.argtypes = [C.c_void_p, _fp, C.POINTER(C.c_uint8), C.c_int32, C.c_int32, _dp, _dp, C.c_double, _dp, C.POINTER(C.c_uint8)]
    lib.gto_plan_cost.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _dp, C.c_int32, _dp, _dp, _dp]
    lib.gto_base_place.argtypes = [C.c_void_p, C.POINTER(BaseIn), C.POINTER(Options), C.POINTER(BaseOut), _dp]
    if path == os.environ.get("GTO_B200_LIB", LIB_PATH):
        _lib = lib
    return lib


def default_options(**overrides) -> Options:
    o = Options()
    load_library().gto_default_options(C.byref(o))
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError(f"unknown solver option {k}")
        setattr(o, k, v)
    return o


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


@dataclass
class Batch:
    """Host-side description of B independent problems (``gto_batch_in``)."""

    T: int
    dt: float
    qc: np.ndarray  # [B,ndof]
    q_seed: np.ndarray  # [B,T,ndof]
    goal_tf: np.ndarray  # [B,2,3,4]
    base_position: Optional[np.ndarray] = None  # [B,3]
    field_all: Optional[np.ndarray] = None  # [B] slots (-1: zero field)
    field_obs: Optional[np.ndarray] = None
    standoff_offset: int = -10
    use_standoff: bool = True
    collision_avoidance: bool = True
    w_goal: float = 1.0
    w_obs: float = 10.0
    w_vel: float = 0.01
    flags: int = 0

    @property
    def B(self) -> int:
        return int(np.asarray(self.qc).shape[0])


def goal_transforms(table: RobotTable, RT: np.ndarray, standoff_distance: float, axis_standoff: str) -> np.ndarray:
    """``RT.G`` and ``RT.S.G`` (reference ``gto/gto_planner.py:93-100``) for a stack of goal poses [B,4,4] -> [B,2,3,4]."""
    RT = np.asarray(RT, dtype=np.float64).reshape(-1, 4, 4)
    G = np.eye(4)
    G[:3] = table.G
    S = np.eye(4)
    S["xyz".index(axis_standoff), 3] = standoff_distance
    out = np.zeros((RT.shape[0], 2, 3, 4))
    out[:, 0] = (RT @ G)[:, :3]
    out[:, 1] = (RT @ S @ G)[:, :3]
    return out


class GtoContext:
    """One solver context bound to one CUDA device."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None):
        self._lib = load_library(lib_path)
        self._h = C.c_void_p()
        rc = self._lib.gto_create(C.byref(self._h), int(device))
        if rc != GTO_OK:
            self._h = C.c_void_p()
            raise GtoError(rc, "gto_create failed (no CUDA device / not sm_100?) -- there is no CPU fallback")
        self.device = device
        self.table: Optional[RobotTable] = None
        self.field_keys = {}  # slot -> caller-defined key of the field uploaded last (see B200Solver.upload_field_cached)
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.gto_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != GTO_OK:
            raise GtoError(rc, self._lib.gto_last_error(self._h).decode())

    # ------------------------------------------------------------------------------------------------
    def set_robot(self, t: RobotTable):
        arrs = dict(
            opt_qidx=_i(t.opt_qidx), lo=_d(t.lo), hi=_d(t.hi), mov_parent=_i(t.mov_parent), mov_type=_i(t.mov_type),
            mov_origin=_d(t.mov_origin), mov_axis=_d(t.mov_axis), mov_qidx=_i(t.mov_qidx), mov_opt=_i(t.mov_opt),
            link_mov=_i(t.link_mov), link_tf=_d(t.link_tf), link_pt_start=_i(t.link_pt_start), link_pt_count=_i(t.link_pt_count),
            link_optmask=np.ascontiguousarray(t.link_optmask, dtype=np.uint32), points=np.ascontiguousarray(t.points, dtype=np.float32),
            grip_tf=_d(t.grip_tf),
        )
        d = RobotDesc()
        d.ndof, d.nopt, d.nmov, d.nlinks, d.npoints = t.ndof, t.nopt, t.nmov, t.nlinks, t.npoints
        for k in ("opt_qidx", "mov_parent", "mov_type", "mov_qidx", "mov_opt", "link_mov", "link_pt_start", "link_pt_count"):
            setattr(d, k, _ptr(arrs[k], _ip))
        for k in ("lo", "hi", "mov_origin", "mov_axis", "link_tf", "grip_tf"):
            setattr(d, k, _ptr(arrs[k], _dp))
        d.link_optmask = _ptr(arrs["link_optmask"], _up)
        d.points = _ptr(arrs["points"], _fp)
        d.grip_mov, d.grip_pt_start, d.grip_pt_count, d.grip_optmask = int(t.grip_mov), int(t.grip_pt_start), int(t.grip_pt_count), int(t.grip_optmask)
        self._check(self._lib.gto_set_robot(self._h, C.byref(d)))
        self.table = t

    def set_field(self, slot: int, cost: np.ndarray, origin: Sequence[float], pitch: float):
        cost = np.ascontiguousarray(cost, dtype=np.float32)
        if cost.ndim != 3:
            raise ValueError("cost field must be [Nx,Ny,Nz]")
        dims = _i(cost.shape)
        org = _d(np.asarray(origin).reshape(3))
        self.field_keys.pop(int(slot), None)
        self._check(self._lib.gto_set_field(self._h, int(slot), _ptr(cost, _fp), _ptr(dims, _ip), _ptr(org, _dp), float(pitch)))

    # ------------------------------------------------------------------------------------------------
    def _batch_in(self, b: Batch):
        t = self.table
        if t is None:
            raise GtoError(-4, "set_robot first")
        B, T = b.B, int(b.T)
        keep = dict(
            qc=_d(b.qc).reshape(B, t.ndof), q_seed=_d(b.q_seed).reshape(B, T, t.ndof), goal_tf=_d(b.goal_tf).reshape(B, 24),
            base=_d(b.base_position if b.base_position is not None else np.zeros((B, 3))).reshape(B, 3),
            fa=_i(b.field_all if b.field_all is not None else -np.ones(B)).reshape(B),
            fo=_i(b.field_obs if b.field_obs is not None else -np.ones(B)).reshape(B),
        )
        s = BatchIn()
        s.B, s.T, s.dt = B, T, float(b.dt)
        s.qc, s.q_seed, s.goal_tf, s.base_position = (_ptr(keep[k], _dp) for k in ("qc", "q_seed", "goal_tf", "base"))
        s.field_all, s.field_obs = _ptr(keep["fa"], _ip), _ptr(keep["fo"], _ip)
        s.standoff_offset, s.use_standoff, s.collision_avoidance = int(b.standoff_offset), int(bool(b.use_standoff)), int(bool(b.collision_avoidance))
        s.w_goal, s.w_obs, s.w_vel, s.flags = float(b.w_goal), float(b.w_obs), float(b.w_vel), int(b.flags)
        return s, keep

    def _batch_out(self, B: int, T: int, out: Optional[dict] = None):
        nd = self.table.ndof
        shapes = dict(Q=((B, T, nd), np.float64), dQ=((B, T - 1, nd), np.float64), cost=((B,), np.float64), iters=((B,), np.int32), status=((B,), np.int32))
        if out is not None:  # caller-owned (e.g. pinned) result buffers, reused across calls
            for k, (shp, dt) in shapes.items():
                a = out.get(k)
                if a is None or a.shape != shp or a.dtype != dt or not a.flags["C_CONTIGUOUS"]:
                    raise ValueError(f"out[{k!r}] must be a C-contiguous {np.dtype(dt).name} array of shape {shp}")
            res = out
        else:
            res = {k: np.empty(shp, dt) for k, (shp, dt) in shapes.items()}
        o = BatchOut()
        o.Q, o.dQ, o.cost = _ptr(res["Q"], _dp), _ptr(res["dQ"], _dp), _ptr(res["cost"], _dp)
        o.iters, o.status = _ptr(res["iters"], _ip), _ptr(res["status"], _ip)
        return o, res

    def solve_batch(self, b: Batch, options: Optional[Options] = None, out: Optional[dict] = None) -> dict:
        s, keep = self._batch_in(b)
        o, res = self._batch_out(b.B, int(b.T), out)
        self._check(self._lib.gto_solve_batch(self._h, C.byref(s), C.byref(options) if options is not None else None, C.byref(o)))
        return res

    def upload_batch(self, b: Batch):
        s, keep = self._batch_in(b)
        self._check(self._lib.gto_upload_batch(self._h, C.byref(s)))
        self._resident = (b.B, int(b.T))

    def solve_resident(self, options: Optional[Options] = None):
        self._check(self._lib.gto_solve_resident(self._h, C.byref(options) if options is not None else None))

    def download_batch(self) -> dict:
        B, T = self._resident
        o, res = self._batch_out(B, T)
        self._check(self._lib.gto_download_batch(self._h, C.byref(o)))
        return res

    def result_device_ptr(self):
        p = C.c_void_p()
        n = C.c_int64()
        self._check(self._lib.gto_result_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def eval_batch(self, b: Batch, want_rows: bool = True) -> dict:
        t = self.table
        s, keep = self._batch_in(b)
        B, T, n = b.B, int(b.T), t.nopt
        nrows = (T * t.npoints if b.collision_avoidance else 0) + 3 * t.grip_pt_count * (2 if b.use_standoff else 1)
        res = dict(H=np.zeros((B, T, n, n), np.float32), g=np.zeros((B, T, n), np.float64), cost=np.zeros((B, T), np.float64))
        e = EvalOut()
        if want_rows:
            res["rows"] = np.zeros((B, nrows, n + 1), np.float32)
            e.rows = _ptr(res["rows"], _fp)
        e.H, e.g, e.cost = _ptr(res["H"], _fp), _ptr(res["g"], _dp), _ptr(res["cost"], _dp)
        self._check(self._lib.gto_eval_batch(self._h, C.byref(s), C.byref(e)))
        return res

    def configure(self, **knobs):
        """Run-time tuning knobs (``gto_configure``): jrows_budget_mb, pdl, launch_events, step_fk, cull_nslot, cons_warps,
        slot_floats, step_dbg (iteration whose step launch reports its phase clocks), fused, blocking_sync (the host thread sleeps
        in the convergence polls instead of spinning: for more solving threads than cores)."""
        for k, v in knobs.items():
            self._check(self._lib.gto_configure(self._h, k.encode(), float(v)))

    def profile(self) -> dict:
        p = Profile()
        self._check(self._lib.gto_get_profile(self._h, C.byref(p)))
        return p.as_dict()

    def cloud_set(self, points: np.ndarray):
        """World-frame points of a depth image (reference ``DepthPointCloud.points``, the KD-tree's data)."""
        pts = _d(points).reshape(-1, 3)
        self._check(self._lib.gto_cloud_set(self._h, _ptr(pts, _dp), pts.shape[0]))

    def cloud_backproject(self, depth: np.ndarray, K: np.ndarray, cam_pose: np.ndarray, threshold: float, target_mask=None) -> np.ndarray:
        """Depth image -> world-frame point cloud [M,3] on the device (reference ``DepthPointCloud.__init__`` / ``backproject_camera``)."""
        dep = np.ascontiguousarray(depth, dtype=np.float32)
        H, W = dep.shape
        Kinv, pose = _d(np.linalg.inv(np.asarray(K, dtype=np.float64)).reshape(9)), _d(np.asarray(cam_pose, dtype=np.float64).reshape(16))
        mask = None if target_mask is None else np.ascontiguousarray(np.asarray(target_mask).reshape(H, W) != 0, dtype=np.uint8)
        pts, valid = np.zeros((H * W, 3)), np.zeros(H * W, np.uint8)
        self._check(self._lib.gto_cloud_backproject(self._h, _ptr(dep, _fp), None if mask is None else mask.ctypes.data_as(C.POINTER(C.c_uint8)), H, W,
                                                    _ptr(Kinv, _dp), _ptr(pose, _dp), float(threshold), _ptr(pts, _dp), valid.ctypes.data_as(C.POINTER(C.c_uint8))))
        return pts[valid != 0]

    def cloud_query(self, query: np.ndarray, depth: np.ndarray, K: np.ndarray, cam_inv: np.ndarray, mode: int, epsilon=0.02, w_inside=1.0):
        """Signed distance (mode 0), cost (mode 1) or visibility (mode 2: 1 = outside) of ``query`` [N,3] w.r.t. the cloud; returns
        (float32 [N], kernel ms)."""
        q = _d(query).reshape(-1, 3)
        dep = np.ascontiguousarray(depth, dtype=np.float32)
        Km, Ri = _d(np.asarray(K).reshape(9)), _d(np.asarray(cam_inv).reshape(16))
        out = np.zeros(q.shape[0], np.float32)
        ms = np.zeros(1)
        self._check(self._lib.gto_cloud_query(self._h, _ptr(q, _dp), q.shape[0], _ptr(dep, _fp), dep.shape[0], dep.shape[1], _ptr(Km, _dp),
                                              _ptr(Ri, _dp), int(mode), float(epsilon), float(w_inside), _ptr(out, _fp), _ptr(ms, _dp)))
        return out, float(ms[0])

    def plan_cost(self, plans: np.ndarray, field_slot: int, base_position=(0.0, 0.0, 0.0)):
        """plans [n,T,ndof] -> (cost [n], dist [n]) (reference ``GTORobotModel.compute_plan_cost``)."""
        plans = _d(plans)
        n, T = plans.shape[0], plans.shape[1]
        cost, dist = np.zeros(n), np.zeros(n)
        bp = _d(np.asarray(base_position).reshape(3))
        self._check(self._lib.gto_plan_cost(self._h, n, T, _ptr(plans, _dp), int(field_slot), _ptr(bp, _dp), _ptr(cost, _dp), _ptr(dist, _dp)))
        return cost, dist

    def base_place(self, qc: np.ndarray, RTs: np.ndarray, w_effort: float = 0.01, occupancy: Optional[np.ndarray] = None,
                   occ_origin=(0.0, 0.0), occ_resolution: float = 0.05, options: Optional[Options] = None) -> dict:
        """Mobile-base placement (reference ``BasePlanner.plan_goalset``, gto/base_planner.py:94-168) for a stack of problems:
        ``RTs`` [B,n,4,4] goal poses in the current base frame -> Q [B,n,ndof], y [B,3], cost/collision/iters/status [B]."""
        t = self.table
        RTs = np.asarray(RTs, dtype=np.float64)
        if RTs.ndim == 3:
            RTs = RTs[None]
        B, n = RTs.shape[0], RTs.shape[1]
        G = np.eye(4)
        G[:3] = t.G
        goal = _d((RTs @ G)[:, :, :3, :])
        qcv = _d(np.asarray(qc, dtype=np.float64).reshape(-1))
        if qcv.shape[0] != t.ndof:
            raise ValueError(f"qc must have {t.ndof} entries")
        bi = BaseIn()
        bi.B, bi.n_goals, bi.qc, bi.goal_tf, bi.w_effort = B, n, _ptr(qcv, _dp), _ptr(goal, _dp), float(w_effort)
        occ = None
        if occupancy is not None:
            occ = np.ascontiguousarray(occupancy, dtype=np.float32)
            if occ.ndim != 2:
                raise ValueError("occupancy must be [nx,ny]")
            bi.occupancy = _ptr(occ, _fp)
            bi.occ_dims[0], bi.occ_dims[1] = occ.shape
            bi.occ_origin[0], bi.occ_origin[1] = float(occ_origin[0]), float(occ_origin[1])
            bi.occ_resolution = float(occ_resolution)
        res = dict(Q=np.zeros((B, n, t.ndof)), y=np.zeros((B, 3)), cost=np.zeros(B), collision=np.zeros(B),
                   iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32))
        bo = BaseOut()
        bo.Q, bo.y, bo.cost, bo.collision = (_ptr(res[k], _dp) for k in ("Q", "y", "cost", "collision"))
        bo.iters, bo.status = _ptr(res["iters"], _ip), _ptr(res["status"], _ip)
        ms = np.zeros(1)
        self._check(self._lib.gto_base_place(self._h, C.byref(bi), C.byref(options) if options is not None else None, C.byref(bo), _ptr(ms, _dp)))
        res["kernel_ms"] = float(ms[0])
        return res
