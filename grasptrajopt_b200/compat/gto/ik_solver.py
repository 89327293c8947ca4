"""``IKSolver`` on the B200 solver (reference ``gto/ik_solver.py:18-115``): the T=1 point-matching problem

    min_q  sum_k | FK_gripper(q) x_k - (RT.G) x_k |^2     s.t.  lo <= q <= hi

is the trajectory problem with one free knot, no velocity term and no stand-off, so it runs on the same kernels
(3 knots: two pinned at the seed, one free).  Many goals can be solved in one batch with ``solve_ik_batch``.

``collision_avoidance=True`` adds the reference's IK collision term ``10 * sum(c)`` -- unsquared (:69), unlike the planner's
``10 * sum(c^2)`` -- through ``GTO_FLAG_OBS_LINEAR`` (value ``w*c``, gradient ``w*dc/dq`` of the trilinear field, no Gauss-Newton
curvature).  In the reference the term is a nearest-node gather with structurally zero gradient and it is
disabled in every shipped experiment (``ik_collision_avoidance=False``, Q14).
"""
from __future__ import annotations

import numpy as np

import optas
from gto.b200_solver import B200Solver
from grasptrajopt_b200 import capi
from grasptrajopt_b200.spatial import mat2quat_wxyz


class IKSolver:
    def __init__(self, robot, link_ee, link_gripper, collision_avoidance=True, device=0):
        self.robot = robot
        self.link_ee = link_ee
        self.link_gripper = link_gripper
        self.robot_name = robot.get_name()
        self.gripper_points = robot.surface_pc_map[link_gripper].points
        self.gripper_tf = robot.get_link_transform_function(link=link_gripper, base_link=link_ee)
        self.collision_avoidance = collision_avoidance
        self.device = device
        self.solver = None

    def setup_optimization(self):
        self.fk = self.robot.get_global_link_transform_function(link=self.link_ee)
        options = capi.default_options(max_iter=50)  # reference: max_iter 50 (:75)
        if not self.collision_avoidance:
            options.bundle = 0  # goal rows only: a smooth problem, no field kinks for the cutting planes to resolve -> plain LM
        self.solver = B200Solver(self.robot, self.link_ee, self.link_gripper, T=3, dt=1.0, use_standoff=False, standoff_offset=-1,
                                 collision_avoidance=self.collision_avoidance, w_vel=0.0, device=self.device, obs_linear=True,
                                 options=options)

    def _errors(self, q, RT):
        tf = self.fk(q).toarray()
        err_pos = np.linalg.norm(RT[:3, 3] - tf[:3, 3])
        quat1, quat2 = mat2quat_wxyz(RT[:3, :3]), mat2quat_wxyz(tf[:3, :3])
        err_rot = np.arccos(np.clip(2 * np.square(np.dot(quat1, quat2)) - 1, -1, 1)) * 180 / np.pi
        return err_pos, err_rot

    def solve_ik(self, q_0, RT, sdf_cost_obstacle, base_position):
        q_0 = np.asarray(q_0, dtype=np.float64).reshape(-1)
        RT = np.asarray(RT, dtype=np.float64)
        Q0 = np.tile(q_0.reshape(-1, 1), (1, 3))
        self.solver.reset_initial_seed({f"{self.robot_name}/q/x": self.robot.extract_optimized_dimensions(Q0)})
        params = {f"{self.robot_name}/q/p": self.robot.extract_parameter_dimensions(Q0), "tf_goal": RT, "qc": q_0}
        if self.collision_avoidance:
            params["sdf_cost_obstacle"] = optas.DM(np.asarray(sdf_cost_obstacle).reshape(-1))
            params["base_position"] = optas.DM(base_position)
        self.solver.reset_parameters(params)
        # knots 0 and 1 are pinned at the seed; the IK unknown is knot 2.  The seed itself must stay free, so the pinned
        # knots only serve as the (costless, w_vel = 0) anchor of the trajectory layout.
        solution = self.solver.solve()
        q = solution[f"{self.robot_name}/q"].toarray()[:, 2]
        err_pos, err_rot = self._errors(q, RT)
        if self.collision_avoidance:
            cost, _ = self.robot.compute_plan_cost(q.reshape(-1, 1), np.asarray(sdf_cost_obstacle).reshape(-1), base_position)
        else:
            cost = 0
        return q.flatten(), err_pos, err_rot, cost

    def solve_ik_batch(self, q_0, RTs, sdf_cost_obstacle=None, base_position=None):
        """All candidate grasps of an object in ONE batch (the reference loops ``solve_ik`` over the grasps,
        ``examples/pybullet_gto_planning.py:243-273``).  ``q_0`` [ndof] (shared seed) or [B,ndof], ``RTs`` [B,4,4].
        Returns ``(q [B,ndof], err_pos [B], err_rot_deg [B], cost [B], status [B])`` -- per goal the same values as ``solve_ik``."""
        from grasptrajopt_b200 import kinematics as K

        t = self.solver.table
        RTs = np.asarray(RTs, dtype=np.float64).reshape(-1, 4, 4)
        B = RTs.shape[0]
        q0 = np.broadcast_to(np.asarray(q_0, dtype=np.float64).reshape(-1, t.ndof), (B, t.ndof))
        seeds = np.repeat(q0[:, None, :], 3, axis=1)
        fo = -1
        if self.collision_avoidance and sdf_cost_obstacle is not None and np.any(sdf_cost_obstacle):
            self.solver.p = {"sdf_cost_obstacle": np.asarray(sdf_cost_obstacle, dtype=np.float64).reshape(-1, 1)}
            from gto.b200_solver import get_context
            fo = self.solver._field(get_context(self.device), self.solver.FIELD_OBS, "sdf_cost_obstacle")
        res = self.solver.solve_many(q0, seeds, RTs, base_position=base_position, field_all=fo, field_obs=fo)
        q = res["Q"][:, 2, :]
        tf = K.ee_frames(t, q)
        err_pos = np.linalg.norm(RTs[:, :3, 3] - tf[:, :3, 3], axis=1)
        err_rot = np.array([np.arccos(np.clip(2 * np.square(np.dot(mat2quat_wxyz(RTs[i, :3, :3]), mat2quat_wxyz(tf[i, :3, :3]))) - 1, -1, 1)) * 180 / np.pi
                            for i in range(B)])
        if fo >= 0:
            cost = np.array([self.robot.compute_plan_cost(q[i].reshape(-1, 1), np.asarray(sdf_cost_obstacle).reshape(-1),
                                                          [0, 0, 0] if base_position is None else base_position)[0] for i in range(B)])
        else:
            cost = np.zeros(B)
        return q, err_pos, err_rot, cost, res["status"]

    def solve_fk(self, q_0):
        return self.fk(q_0).toarray()
