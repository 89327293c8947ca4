"""``B200Solver``: the solver object behind ``GTOPlanner.solver`` / ``IKSolver.solver``.

It replaces, for the grasp-trajectory problem family, the reference's ``optas.CasADiSolver`` (``optas/solver.py:323-421``,
driven from ``gto/gto_planner.py:142,160-176,222-239``) and keeps its protocol:

    solver.reset_initial_seed({f"{robot}/q/x": nopt-by-T})
    solver.reset_parameters({"qc", "tf_goal" (16-by-n), "sdf_cost_all", "sdf_cost_obstacle", "base_position", f"{robot}/q/p"})
    solution = solver.solve()      ->  {f"{robot}/q": DM ndof-by-T, f"{robot}/dq": DM ndof-by-(T-1), "f": DM 1x1, ...}
    solver.stats() / did_solve() / number_of_iterations()

``tf_goal`` with n > 1 columns is the reference's goal set (``mmin`` over per-goal costs, ``gto_planner.py:91-105``): the n
goals are solved as n independent problems in ONE batch on the GPU and the cheapest one is returned -- the same global
problem, since min over Q of min over goals equals min over goals of min over Q.  All per-goal solutions stay
available in ``solver.batch_result``.

Everything numeric happens in ``libgto_b200.so`` (CUDA); there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from optas.dm import DM, _as2d
from grasptrajopt_b200 import capi
from grasptrajopt_b200.goalset import goalset_best

_CONTEXTS: Dict[int, capi.GtoContext] = {}


def get_context(device: int = 0) -> capi.GtoContext:
    """One solver context per (process, device), shared by all planners."""
    ctx = _CONTEXTS.get(device)
    if ctx is None:
        ctx = capi.GtoContext(device)
        _CONTEXTS[device] = ctx
    return ctx


class B200Solver:
    FIELD_ALL, FIELD_OBS = 0, 1  # field slots used by the planner

    def __init__(self, robot, link_ee, link_gripper, T, dt, *, standoff_distance=-0.1, standoff_offset=-10, use_standoff=False,
                 axis_standoff="x", collision_avoidance=True, w_goal=1.0, w_obs=10.0, w_vel=0.01, device=0, options=None, obs_linear=False):
        self.robot = robot
        self.robot_name = robot.get_name()
        self.table = robot.to_table(link_ee, link_gripper)
        self.T, self.dt = int(T), float(dt)
        self.standoff_distance, self.standoff_offset = float(standoff_distance), int(standoff_offset)
        self.use_standoff, self.axis_standoff = bool(use_standoff), axis_standoff
        self.collision_avoidance = bool(collision_avoidance)
        self.w_goal, self.w_obs, self.w_vel = w_goal, w_obs, w_vel
        self.flags = capi.FLAG_OBS_LINEAR if obs_linear else 0  # unsquared obstacle term of the IK solver (gto/ik_solver.py:69)
        self.device = device
        self.options = options
        self.x0: Optional[np.ndarray] = None
        self.p: Dict[str, np.ndarray] = {}
        self._stats = {"success": False, "iter_count": 0}
        self.batch_result = None

    # one upload per distinct field: the planner re-sends both cost fields as parameters on every call (reference protocol), but
    # gto_set_field (device copy, ~340 tensor maps, summed-volume table) costs more than a solve
    @staticmethod
    def upload_field_cached(ctx, device, slot, cost32, origin, pitch):
        """``ctx.set_field`` unless the same field (shape, geometry and a content checksum) already sits in ``slot``."""
        flat = cost32.reshape(-1)
        key = (cost32.shape, tuple(float(v) for v in origin), float(pitch), float(flat.sum(dtype=np.float64)), float(flat[::7].sum(dtype=np.float64)),
               float(np.dot(flat[::13], np.arange(flat[::13].size, dtype=np.float32) % 251)))
        if ctx.field_keys.get(int(slot)) == key:
            return
        ctx.set_field(slot, cost32, origin, pitch)
        ctx.field_keys[int(slot)] = key

    # -- CasADiSolver protocol ----------------------------------------------------------------------------------
    def setup(self, solver_name: str = "b200", solver_options: Optional[dict] = None):
        """Kept for call compatibility (``CasADiSolver(...).setup("ipopt", solver_options=...)``).  ``max_iter`` of an
        ``{'ipopt': {...}}`` dictionary is honoured; IPOPT's ``tol`` has no counterpart (the reference's 1e-15 is unreachable, Q5)."""
        if solver_options:
            ip = solver_options.get("ipopt", solver_options)
            if "max_iter" in ip:
                self.options = capi.default_options(max_iter=int(ip["max_iter"]))
        return self

    def reset_initial_seed(self, x0: Dict[str, np.ndarray]) -> None:
        key = f"{self.robot_name}/q/x"
        if key in x0:
            self.x0 = np.asarray(_as2d(x0[key]), dtype=np.float64)  # nopt-by-T
        # dq/x seeds are accepted and ignored: dQ is eliminated analytically (gto_planner.py:68-72)

    def reset_parameters(self, p: Dict[str, np.ndarray]) -> None:
        self.p = {k: np.asarray(_as2d(v), dtype=np.float64) for k, v in p.items()}

    def _field(self, ctx, slot, key):
        if key not in self.p:
            return -1  # missing parameter == zeros (optas/mx_container.py:121): plan() never sets sdf_cost_all (Q4)
        cost = self.p[key].reshape(-1)
        shape = tuple(int(s) for s in self.robot.field_shape)
        if cost.size != int(np.prod(shape)):
            raise ValueError(f"{key} has {cost.size} entries, the robot's field has {int(np.prod(shape))}")
        if not np.any(cost):
            return -1
        self.upload_field_cached(ctx, self.device, slot, cost.reshape(shape).astype(np.float32), np.asarray(self.robot.origin).reshape(3),
                                 float(self.robot.grid_resolution))
        return slot

    def solve(self) -> Dict[str, DM]:
        t, T, nd = self.table, self.T, self.table.ndof
        ctx = get_context(self.device)
        if ctx.table is not t:
            ctx.set_robot(t)
        tf_goal = self.p["tf_goal"]
        if tf_goal.shape == (4, 4):  # IKSolver passes the 4x4 pose directly (gto/ik_solver.py:37,88)
            RT = tf_goal.reshape(1, 4, 4)
        else:  # 16-by-n, column i = row-major flatten of RT_i (gto_planner.py:148,191)
            RT = tf_goal.T.reshape(-1, 4, 4)
        n = RT.shape[0]
        qc = self.p["qc"].reshape(-1) if "qc" in self.p else None
        Qp = self.p.get(f"{self.robot_name}/q/p")  # npar-by-T
        seed = np.zeros((T, nd))
        if self.x0 is not None:
            seed[:, t.opt_qidx] = self.x0.T
        if Qp is not None and t.par_qidx.size:
            seed[:, t.par_qidx] = Qp.T
        if qc is None:
            qc = seed[0].copy()
        fa = self._field(ctx, self.FIELD_ALL, "sdf_cost_all") if self.collision_avoidance else -1
        fo = self._field(ctx, self.FIELD_OBS, "sdf_cost_obstacle") if self.collision_avoidance else -1
        base = self.p.get("base_position", np.zeros(3)).reshape(-1)[:3]
        batch = capi.Batch(
            T=T, dt=self.dt, qc=np.tile(qc, (n, 1)), q_seed=np.tile(seed, (n, 1, 1)),
            goal_tf=capi.goal_transforms(t, RT, self.standoff_distance, self.axis_standoff),
            base_position=np.tile(base, (n, 1)), field_all=np.full(n, fa, np.int32), field_obs=np.full(n, fo, np.int32),
            standoff_offset=self.standoff_offset, use_standoff=self.use_standoff,
            collision_avoidance=self.collision_avoidance and (fa >= 0 or fo >= 0),
            w_goal=self.w_goal, w_obs=self.w_obs, w_vel=self.w_vel, flags=self.flags)
        res = ctx.solve_batch(batch, self.options)
        self.batch_result = res
        best = goalset_best(res["cost"], res["status"])  # converged goals first, never a NaN solve (grasptrajopt_b200/goalset.py)
        self._stats = {"success": bool(res["status"][best] == capi.STATUS_CONVERGED), "iter_count": int(res["iters"][best]),
                       "status": int(res["status"][best]), "best_goal": best, "profile": ctx.profile()}
        Q, dQ = res["Q"][best].T, res["dQ"][best].T  # ndof-by-T, ndof-by-(T-1)
        name = self.robot_name
        return {f"{name}/q": DM(Q), f"{name}/dq": DM(dQ), f"{name}/q/x": DM(Q[t.opt_qidx]), f"{name}/dq/x": DM(dQ[t.opt_qidx]),
                "f": DM(np.array([[res["cost"][best]]]))}

    def solve_many(self, qc, q_seed, RT, base_position=None, field_all=-1, field_obs=-1) -> dict:
        """B independent problems in ONE batch (no arg-min): ``qc`` [B,ndof], ``q_seed`` [B,T,ndof], ``RT`` [B,4,4], optional
        ``base_position`` [B,3] or [3]; ``field_all`` / ``field_obs``: field slots already uploaded with ``ctx.set_field`` (or -1).
        Returns the raw per-problem result of ``GtoContext.solve_batch`` (Q [B,T,ndof], dQ, cost, iters, status)."""
        t = self.table
        ctx = get_context(self.device)
        if ctx.table is not t:
            ctx.set_robot(t)
        RT = np.asarray(RT, dtype=np.float64).reshape(-1, 4, 4)
        n = RT.shape[0]
        base = np.zeros((n, 3)) if base_position is None else np.broadcast_to(np.asarray(base_position, dtype=np.float64).reshape(-1, 3), (n, 3))
        collide = self.collision_avoidance and (field_all >= 0 or field_obs >= 0)
        batch = capi.Batch(
            T=self.T, dt=self.dt, qc=np.asarray(qc, dtype=np.float64).reshape(n, t.ndof), q_seed=np.asarray(q_seed, dtype=np.float64).reshape(n, self.T, t.ndof),
            goal_tf=capi.goal_transforms(t, RT, self.standoff_distance, self.axis_standoff), base_position=np.ascontiguousarray(base),
            field_all=np.full(n, field_all, np.int32), field_obs=np.full(n, field_obs, np.int32), standoff_offset=self.standoff_offset,
            use_standoff=self.use_standoff, collision_avoidance=collide, w_goal=self.w_goal, w_obs=self.w_obs, w_vel=self.w_vel, flags=self.flags)
        res = ctx.solve_batch(batch, self.options)
        self.batch_result = res
        res["profile"] = ctx.profile()
        return res

    def stats(self) -> dict:
        return self._stats

    def did_solve(self) -> bool:
        return bool(self._stats["success"])

    def number_of_iterations(self) -> int:
        return int(self._stats["iter_count"])
