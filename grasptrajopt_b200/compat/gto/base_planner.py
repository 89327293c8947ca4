"""``BasePlanner`` on the B200 solver: constructor, ``setup_optimization(goal_size, base_effort_weight)`` and
``plan_goalset(qc, RTs) -> (Q [ndof x n], y [3], err_pos [n], err_rot [n], collision cost)`` as in the reference
(``gto/base_planner.py:19-168``).

What changed underneath: no symbolic NLP is built; the problem goes through the C-ABI entry ``gto_base_place`` (kernel
``k_base_place``, ``grasptrajopt_b200/csrc/base_place.cuh``).  The reference calls ``plan_goalset`` in a rejection loop over random
grasp subsets until the occupancy-grid collision count is zero (``examples/pybullet_gto_planning_mobile.py:187-201``);
``plan_goalset_batch`` solves all candidate subsets in one launch and ``plan_until_collision_free`` is that loop on top of it.
Everything numeric happens in ``libgto_b200.so`` (CUDA); there is no CPU path.
"""
from __future__ import annotations

import numpy as np

from optas.models import TaskModel
from gto.b200_solver import get_context
from gto.utils import rotZ
from grasptrajopt_b200 import capi
from grasptrajopt_b200.spatial import mat2quat_wxyz


class BasePlanner:
    def __init__(self, robot, link_ee, link_gripper, device=0):
        self.task = TaskModel("base_pose_estimator", dim=3)  # (x, y, theta)
        self.task_name = self.task.name
        self.robot = robot
        self.robot_name = robot.get_name()
        self.link_ee = link_ee
        self.link_gripper = link_gripper
        self.gripper_points = robot.surface_pc_map[link_gripper].points
        self.gripper_tf = robot.get_link_transform_function(link=link_gripper, base_link=link_ee)
        self.device = device
        self.goal_size = None
        self.base_effort_weight = 0.01
        self.options = None
        self.last_result = None

    def setup_optimization(self, goal_size=1, base_effort_weight=0.01):
        if not 1 <= int(goal_size) <= 32:
            raise ValueError("BasePlanner on the B200 solver takes 1..32 goals per problem (one warp lane per goal)")
        self.goal_size = int(goal_size)
        self.base_effort_weight = float(base_effort_weight)
        self.fk = self.robot.get_global_link_transform_function(self.link_gripper, n=self.goal_size)
        self.options = capi.default_options(max_iter=100)  # reference: {'ipopt': {'max_iter': 100, 'tol': 1e-15}} (:92)
        self.table = self.robot.to_table(self.link_ee, self.link_gripper)

    # -- the batched call ---------------------------------------------------------------------------------------------------
    def _occupancy(self):
        r = self.robot
        if getattr(r, "occupancy_grid", None) is None:
            return None, (0.0, 0.0), 0.05
        grid = np.asarray(r.occupancy_grid, dtype=np.float32).reshape(r.occupancy_grid_shape)
        return grid, np.asarray(r.occupancy_grid_origin).reshape(2), float(r.grid_resolution)

    def plan_goalset_batch(self, qc, RTs_batch):
        """``RTs_batch`` [B,n,4,4] -> dict(Q [B,n,ndof], y [B,3], cost, collision, iters, status [B], err_pos, err_rot [B,n])."""
        if self.goal_size is None:
            raise RuntimeError("call setup_optimization() first")
        RTs_batch = np.asarray(RTs_batch, dtype=np.float64)
        if RTs_batch.ndim != 4 or RTs_batch.shape[1] != self.goal_size:
            raise ValueError(f"expected [B,{self.goal_size},4,4] goal poses")
        qc = np.asarray(qc, dtype=np.float64).reshape(-1)
        ctx = get_context(self.device)
        if ctx.table is not self.table:
            ctx.set_robot(self.table)
        grid, origin, res = self._occupancy()
        out = ctx.base_place(qc, RTs_batch, self.base_effort_weight, grid, origin, res, self.options)
        out["err_pos"], out["err_rot"] = self._errors(out["Q"], out["y"], RTs_batch)
        self.last_result = out
        return out

    def _errors(self, Q, y, RTs):
        """Position / rotation error of every goal in the new base frame (reference :131-148)."""
        B, n = RTs.shape[:2]
        err_pos, err_rot = np.zeros((B, n), np.float32), np.zeros((B, n), np.float32)
        for b in range(B):
            RT_base = rotZ(y[b, 2])
            RT_base[0, 3], RT_base[1, 3] = y[b, 0], y[b, 1]
            for i in range(n):
                G = np.asarray(self.gripper_tf(Q[b, i]).toarray())
                RT = RT_base @ RTs[b, i] @ G
                tf = np.asarray(self.robot.get_global_link_transform(self.link_gripper, Q[b, i]).toarray())
                err_pos[b, i] = np.linalg.norm(RT[:3, 3] - tf[:3, 3])
                q1, q2 = mat2quat_wxyz(RT[:3, :3]), mat2quat_wxyz(tf[:3, :3])
                err_rot[b, i] = np.arccos(np.clip(2 * np.square(np.dot(q1, q2)) - 1, -1, 1)) * 180 / np.pi
        return err_pos, err_rot

    # -- reference signature ------------------------------------------------------------------------------------------------
    def plan_goalset(self, qc, RTs):
        RTs = np.asarray(RTs, dtype=np.float64)
        out = self.plan_goalset_batch(qc, RTs[None])
        return out["Q"][0].T.copy(), out["y"][0].copy(), out["err_pos"][0], out["err_rot"][0], float(out["collision"][0])

    def plan_until_collision_free(self, qc, RTs_per_object, num=2, batch=64, rng=None, max_rounds=16):
        """The reference's loop (examples/pybullet_gto_planning_mobile.py:187-201): draw ``num`` grasps per object, plan the base,
        repeat until the robot at the new base touches no occupied cell -- ``batch`` draws per launch.  Returns the first
        collision-free draw as ``(Q [ndof x n], y, err_pos, err_rot, cost, RTs_all)`` or the cheapest one seen."""
        rng = rng or np.random.default_rng()
        best = None
        for _ in range(max_rounds):
            draws = np.stack([np.concatenate([RT[rng.choice(RT.shape[0], num)] for RT in RTs_per_object if RT.shape[0] > 0])
                              for _ in range(batch)])
            out = self.plan_goalset_batch(qc, draws)
            free = np.flatnonzero(out["collision"] == 0)
            b = int(free[0]) if free.size else int(np.argmin(out["collision"]))
            cand = (out["Q"][b].T.copy(), out["y"][b].copy(), out["err_pos"][b], out["err_rot"][b], float(out["collision"][b]), draws[b])
            if best is None or cand[4] < best[4]:
                best = cand
            if free.size:
                break
        return best
