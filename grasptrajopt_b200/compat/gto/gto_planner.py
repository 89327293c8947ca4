"""``GTOPlanner`` on the B200 solver: same constructor, attributes and ``setup_optimization`` / ``plan`` / ``plan_goalset``
signatures and return types as the reference (``gto/gto_planner.py:21-245``).

What changed underneath: ``setup_optimization`` no longer builds a symbolic CasADi NLP on every call (Q3) -- it only creates
a ``B200Solver`` holding the flags; ``plan_goalset`` solves the n goals as one GPU batch and ranks the interpolated seeds
with a CUDA value-only pass instead of T NumPy FK sweeps per candidate.
"""
from __future__ import annotations

import numpy as np

import optas
from gto.b200_solver import B200Solver, get_context
from gto.utils import interpolate_waypoints


class GTOPlanner:
    def __init__(self, robot, link_ee, link_gripper, collision_avoidance=True, standoff_distance=-0.1, standoff_offset=-10, device=0):
        self._T = 50  # no. time steps in trajectory
        self.Tmax = 10.0
        self.dt = self.Tmax / (self._T - 1)
        self.standoff_offset = standoff_offset
        self.standoff_distance = standoff_distance
        self.robot = robot
        self.robot_name = robot.get_name()
        self.link_ee = link_ee
        self.link_gripper = link_gripper
        self.gripper_points = robot.surface_pc_map[link_gripper].points
        self.gripper_tf = robot.get_link_transform_function(link=link_gripper, base_link=link_ee)
        self.collision_avoidance = collision_avoidance
        self.device = device
        self.solver = None

    # ``planner.T = 30`` behaves as in the reference: T is a plain attribute there and dt is computed once in __init__ from T = 50
    # (gto/gto_planner.py:25-28), so a changed T keeps dt = 10/49.  ``set_horizon`` is the explicit way to rescale dt with T.
    @property
    def T(self):
        return self._T

    @T.setter
    def T(self, value):
        self._T = int(value)

    def set_horizon(self, T, rescale_dt=True):
        """Change the number of knots; ``rescale_dt`` also sets dt = Tmax / (T - 1) (what BASELINE's 30-knot configurations use)."""
        self._T = int(T)
        if rescale_dt:
            self.dt = self.Tmax / (self._T - 1)

    def setup_optimization(self, goal_size=1, use_standoff=False, axis_standoff="x"):
        self.fk = self.robot.get_global_link_transform_function(self.link_gripper, n=self.T)
        self.pose_standoff = optas.spatialmath.standoff(self.standoff_distance, axis_standoff)
        self.solver = B200Solver(
            self.robot, self.link_ee, self.link_gripper, self.T, self.dt, standoff_distance=self.standoff_distance,
            standoff_offset=self.standoff_offset, use_standoff=use_standoff, axis_standoff=axis_standoff,
            collision_avoidance=self.collision_avoidance, device=self.device,
        ).setup("b200", solver_options={"ipopt": {"max_iter": 100, "tol": 1e-15}})

    def _seed_from(self, qc, q_solution):
        data = interpolate_waypoints(np.stack([qc, q_solution]), self.T, self.robot.ndof)
        index = np.array(self.robot.parameter_joint_indexes).astype(np.int32)
        data[:, index] = np.array(qc)[index]
        return data.T  # ndof-by-T

    def _solve(self, qc, tf_goal, Q0, sdf_cost_all, sdf_cost_obstacle, base_position):
        Q0 = optas.DM(Q0)
        self.solver.reset_initial_seed({f"{self.robot_name}/q/x": self.robot.extract_optimized_dimensions(Q0)})
        params = {
            "qc": optas.DM(qc),
            "tf_goal": optas.DM(tf_goal),
            "sdf_cost_obstacle": optas.DM(np.asarray(sdf_cost_obstacle).reshape(-1)),
            "base_position": optas.DM(base_position),
            f"{self.robot_name}/q/p": self.robot.extract_parameter_dimensions(Q0),
        }
        if sdf_cost_all is not None:
            params["sdf_cost_all"] = optas.DM(np.asarray(sdf_cost_all).reshape(-1))
        self.solver.reset_parameters(params)
        solution = self.solver.solve()
        Q = solution[f"{self.robot_name}/q"]
        dQ = solution[f"{self.robot_name}/dq"]
        return Q.toarray(), dQ.toarray(), solution["f"].toarray().flatten()

    def plan(self, qc, RT, sdf_cost_obstacle, base_position, q_solution=None, use_standoff=True, axis_standoff="x"):
        self.setup_optimization(goal_size=1, use_standoff=use_standoff, axis_standoff=axis_standoff)
        qc = np.asarray(qc, dtype=np.float64).reshape(-1)
        tf_goal = np.zeros((16, 1))
        tf_goal[:, 0] = np.asarray(RT).flatten()
        if q_solution is None:
            Q0 = np.diag(qc) @ np.ones((self.robot.ndof, self.T))
        else:
            Q0 = self._seed_from(qc, np.asarray(q_solution).reshape(-1))
        # plan() never passes sdf_cost_all: knots before the stand-off knot see a zero field (reference :165-173, Q4)
        return self._solve(qc, tf_goal, Q0, None, sdf_cost_obstacle, base_position)

    def plan_goalset(self, qc, RTs, sdf_cost_all, sdf_cost_obstacle, base_position, q_solutions=None, use_standoff=True,
                     axis_standoff="x", interpolate=True):
        RTs = np.asarray(RTs)
        n = RTs.shape[0]
        self.setup_optimization(goal_size=n, use_standoff=use_standoff, axis_standoff=axis_standoff)
        qc = np.asarray(qc, dtype=np.float64).reshape(-1)
        tf_goal = np.zeros((16, n))
        for i in range(n):
            tf_goal[:, i] = RTs[i].flatten()
        if q_solutions is None:
            Q0 = np.diag(qc) @ np.ones((self.robot.ndof, self.T))
        else:
            q_solutions = np.asarray(q_solutions, dtype=np.float64)
            plans = np.stack([self._seed_from(qc, q_solutions[:, i]) for i in range(q_solutions.shape[1])])  # [k, ndof, T]
            cost_all, dist_all = self._rank_seeds(plans, sdf_cost_obstacle, base_position)
            for i, (c, d) in enumerate(zip(cost_all, dist_all)):
                print(f"plan {i}, cost {c:.2f}, dist {d:.2f}")
            ind = np.lexsort((dist_all, cost_all))  # sort by cost, then by distance
            print("intialize with solution", ind[0])
            if interpolate:
                Q0 = plans[ind[0]]
            else:
                Q0 = np.diag(qc) @ np.ones((self.robot.ndof, self.T))
                for i in range(self.T + self.standoff_offset, self.T):
                    Q0[:, i] = plans[ind[0]][:, self.T - 1]
        return self._solve(qc, tf_goal, Q0, sdf_cost_all, sdf_cost_obstacle, base_position)

    def _rank_seeds(self, plans, sdf_cost_obstacle, base_position):
        """``GTORobotModel.compute_plan_cost`` for all candidates at once on the GPU (reference :204-215, one NumPy FK sweep
        per knot and candidate)."""
        ctx = get_context(self.device)
        table = self.robot.to_table(self.link_ee, self.link_gripper)
        if ctx.table is not table:
            ctx.set_robot(table)
        shape = tuple(int(s) for s in self.robot.field_shape)
        B200Solver.upload_field_cached(ctx, self.device, B200Solver.FIELD_OBS, np.asarray(sdf_cost_obstacle, dtype=np.float32).reshape(shape),
                                       np.asarray(self.robot.origin).reshape(3), float(self.robot.grid_resolution))
        return ctx.plan_cost(np.transpose(plans, (0, 2, 1)), B200Solver.FIELD_OBS, base_position)
