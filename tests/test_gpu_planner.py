"""The reference-facing planner classes (gto.GTOPlanner / gto.IKSolver on B200Solver) end to end on the GPU, checked against
the oracle on the same inputs.  Uses a small synthetic URDF + box meshes written at test time (the GPU box has no
reference checkout).  Run with -m gpu."""
import numpy as np
import pytest

import gto_oracle as O
from grasptrajopt_b200 import capi
from gto.gto_models import GTORobotModel
from gto.gto_planner import GTOPlanner
from gto.ik_solver import IKSolver
from grasptrajopt_b200 import scenes as S

from test_compat_api import _write_box

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def robot(tmp_path_factory):
    d = tmp_path_factory.mktemp("arm3")
    _write_box(d)
    r = GTORobotModel(str(d), urdf_filename=str(d / "arm3.urdf"), time_derivs=[0, 1], param_joints=["slide"],
                      collision_link_names=["base", "l1", "l2", "tool"], sample_point_count=64, seed=1)
    r.setup_workspace_field(arm_len=0.6, arm_height=0.2)
    return r


def _problem(robot, planner, qc, RT, seed, field_all, field_obs, use_standoff, axis):
    t = robot.to_table("tool", "tool")
    shape = tuple(robot.field_shape)
    mk = lambda c: None if c is None or not np.any(c) else O.Field(np.asarray(c, np.float32).reshape(shape), robot.origin.reshape(3), robot.grid_resolution)
    return O.Problem(table=t, T=planner.T, dt=planner.dt, qc=qc, RT=RT, q_seed=seed.T, base_position=np.zeros(3), field_all=mk(field_all),
                     field_obs=mk(field_obs), standoff_offset=planner.standoff_offset, standoff_distance=planner.standoff_distance,
                     axis_standoff=axis, use_standoff=use_standoff, collision_avoidance=True)


def test_plan_matches_oracle(robot):
    planner = GTOPlanner(robot, "tool", "tool", standoff_distance=-0.05)
    planner.T = 30
    assert planner.dt == pytest.approx(10.0 / 49)  # reference semantics: dt is fixed at construction (gto_planner.py:25-28)
    planner.set_horizon(30)
    assert planner.dt == pytest.approx(10.0 / 29)
    qc = np.array([0.2, -0.3, 0.01])
    q_star = np.array([1.1, 0.6, 0.01])
    RT = robot.get_global_link_transform("tool", q_star).toarray()
    sdf = np.zeros(robot.field_size)
    Q, dQ, cost = planner.plan(qc, RT, sdf, [0, 0, 0], q_solution=q_star + 0.05, use_standoff=True, axis_standoff="x")
    assert Q.shape == (3, 30) and dQ.shape == (3, 29) and cost.shape == (1,)
    seed = planner._seed_from(qc, q_star + 0.05)
    r = O.solve_lm(_problem(robot, planner, qc, RT, seed, None, None, True, "x"))
    assert np.abs(Q.T - r.Q).max() < 1e-4
    assert cost[0] == pytest.approx(r.cost, rel=1e-5)
    np.testing.assert_array_equal(Q[2], qc[2])  # parameter joint row constant
    assert planner.solver.did_solve() and planner.solver.number_of_iterations() == r.iters


def test_plan_goalset_returns_cheapest_goal(robot):
    planner = GTOPlanner(robot, "tool", "tool", standoff_distance=-0.05)
    planner.set_horizon(30)
    qc = np.array([0.2, -0.3, 0.01])
    q_goals = np.array([[1.1, 0.6, 0.01], [-1.4, 1.3, 0.01], [0.4, -0.9, 0.01]])
    RTs = np.stack([robot.get_global_link_transform("tool", q).toarray() for q in q_goals])
    # obstacle field: a box near the second goal makes it expensive
    boxes = [((float(RTs[1][0, 3]), float(RTs[1][1, 3]), float(RTs[1][2, 3])), (0.15, 0.15, 0.15))]
    cost = S.sdf_cost(S.box_sdf(robot.workspace_points, boxes)).astype(np.float64)
    Q, dQ, f = planner.plan_goalset(qc, RTs, cost, cost, [0, 0, 0], q_solutions=q_goals.T, use_standoff=True, axis_standoff="x")
    res = planner.solver.batch_result
    assert res["cost"].shape == (3,) and f[0] == pytest.approx(res["cost"].min())
    best = int(np.argmin(res["cost"]))
    assert best != 1
    np.testing.assert_allclose(Q.T, res["Q"][best])
    # the seed chosen by the GPU ranking pass equals the reference's NumPy ranking
    plans = np.stack([planner._seed_from(qc, q) for q in q_goals])
    c_gpu, d_gpu = planner._rank_seeds(plans, cost, [0, 0, 0])
    for i in range(3):
        c_ref, d_ref = robot.compute_plan_cost(plans[i], cost, [0, 0, 0])
        assert c_gpu[i] == pytest.approx(c_ref, rel=1e-4, abs=1e-6) and d_gpu[i] == pytest.approx(d_ref)
    # per-goal solutions agree with the oracle
    seed = plans[np.lexsort((d_gpu, c_gpu))[0]]
    for i in (0, 2):
        r = O.solve_lm(_problem(robot, planner, qc, RTs[i], seed, cost, cost, True, "x"))
        assert res["cost"][i] == pytest.approx(r.cost, rel=1e-3)


def test_ik_solver_reaches_pose(robot):
    ik = IKSolver(robot, "tool", "tool", collision_avoidance=False)
    ik.setup_optimization()
    q_star = np.array([0.9, 0.5, 0.02])
    RT = robot.get_global_link_transform("tool", q_star).toarray()
    q, err_pos, err_rot, cost = ik.solve_ik(np.array([0.5, 0.1, 0.02]).reshape(-1, 1), RT, np.zeros(robot.field_size), [0, 0, 0])
    assert q.shape == (3,) and err_pos < 1e-4 and err_rot < 0.05 and cost == 0
    assert q[2] == 0.02


def test_ik_batch_equals_single_solves(robot):
    """all candidate grasps in one batch == one ``solve_ik`` per grasp (reference loop,
    examples/pybullet_gto_planning.py:243-273), and each reaches its pose."""
    ik = IKSolver(robot, "tool", "tool", collision_avoidance=False)
    ik.setup_optimization()
    rng = np.random.default_rng(5)
    q_stars = np.stack([rng.uniform([-1.2, -1.0, 0.02], [1.2, 1.0, 0.02]) for _ in range(16)])
    RTs = np.stack([robot.get_global_link_transform("tool", q).toarray() for q in q_stars])
    q0 = np.array([0.3, 0.2, 0.02])
    qb, ep, er, cost, status = ik.solve_ik_batch(q0, RTs)
    assert qb.shape == (16, 3) and np.all(status == 0) and np.all(ep < 1e-4) and np.all(er < 0.05) and np.all(cost == 0)
    for i in (0, 7, 15):
        q1, ep1, er1, _ = ik.solve_ik(q0.reshape(-1, 1), RTs[i], np.zeros(robot.field_size), [0, 0, 0])
        np.testing.assert_array_equal(q1, qb[i])
        assert ep1 == pytest.approx(ep[i], abs=1e-12)


def test_base_planner_matches_oracle(robot):
    """gto.BasePlanner (reference gto/base_planner.py:19-168) through the reference signature, against oracle/base_oracle.py."""
    import base_oracle as BO
    from gto.base_planner import BasePlanner

    rng = np.random.default_rng(5)
    robot.setup_occupancy_grid(np.column_stack([rng.uniform(0, 1.2, 400), rng.uniform(-0.8, 0.8, 400), rng.uniform(0.02, 0.5, 400)]))
    bp = BasePlanner(robot, "tool", "tool")
    bp.setup_optimization(goal_size=2, base_effort_weight=0.01)
    qc = np.array([0.2, -0.3, 0.01])
    Tb = BO.base_tf(np.array([0.3, -0.2, 0.25]))
    RTs = np.stack([np.linalg.inv(Tb) @ robot.get_global_link_transform("tool", q).toarray() for q in ([0.9, 0.4, 0.01], [-0.5, 1.0, 0.01])])
    Q, y, err_pos, err_rot, cost = bp.plan_goalset(qc, RTs)
    assert Q.shape == (3, 2) and y.shape == (3,) and err_pos.shape == (2,) and err_rot.shape == (2,)
    t = robot.to_table("tool", "tool")
    grid = np.asarray(robot.occupancy_grid, np.float32).reshape(robot.occupancy_grid_shape)
    r = BO.solve_base(BO.BaseProblem(t, qc, RTs, 0.01, grid, np.asarray(robot.occupancy_grid_origin).reshape(2), robot.grid_resolution))
    tol = 1e-6 if r.status == 0 else 1e-4
    assert np.abs(y - r.y).max() < tol and np.abs(Q.T - r.Q).max() < tol
    assert cost == r.collision
    assert np.allclose(err_pos, r.err_pos, atol=1e-5) and np.allclose(err_rot, r.err_rot, atol=1e-2)
    # the batched form returns the same bits for the same problem
    out = bp.plan_goalset_batch(qc, np.stack([RTs, RTs, RTs]))
    assert np.array_equal(out["y"][0], y) and np.array_equal(out["y"][2], y)


def test_base_planner_rejection_loop(robot):
    """The mobile example's loop (examples/pybullet_gto_planning_mobile.py:187-201) with a batch of draws per launch."""
    import base_oracle as BO
    from gto.base_planner import BasePlanner

    rng = np.random.default_rng(9)
    robot.setup_occupancy_grid(np.column_stack([rng.uniform(0.6, 1.2, 300), rng.uniform(-0.8, 0.8, 300), rng.uniform(0.02, 0.5, 300)]))
    bp = BasePlanner(robot, "tool", "tool")
    bp.setup_optimization(goal_size=2)
    qc = np.array([0.2, -0.3, 0.01])
    Tbi = np.linalg.inv(BO.base_tf(np.array([0.2, 0.1, 0.2])))
    objs = [np.stack([Tbi @ robot.get_global_link_transform("tool", [a, b, 0.01]).toarray() for a in (0.3, 0.9, 1.4)]) for b in (0.4, 1.0)]
    Q, y, err_pos, err_rot, cost, draw = bp.plan_until_collision_free(qc, objs, num=1, batch=8, rng=np.random.default_rng(0))
    assert Q.shape == (3, 2) and y.shape == (3,) and draw.shape == (2, 4, 4) and cost >= 0
    res = bp.last_result
    assert res["y"].shape == (8, 3) and cost == res["collision"].min()
    # the returned draw, solved alone through the reference signature, gives the same placement
    Q1, y1, _, _, c1 = bp.plan_goalset(qc, draw)
    assert np.array_equal(y1, y) and np.array_equal(Q1, Q) and c1 == cost


def _ik_batch_from_workload(w, n, linear_obstacle):
    """The batch ``IKSolver.solve_ik_batch`` hands to the solver (gto/ik_solver.py:30-76 as a 3-knot trajectory: two knots pinned
    at the seed, one free, no velocity term, no stand-off), for the first n grasps of a workload."""
    import dataclasses
    from grasptrajopt_b200 import workloads as W

    b = W.slice_batch(w.batch, 0, n)
    seeds = np.repeat(b.qc[:, None, :], 3, axis=1)
    return dataclasses.replace(b, T=3, dt=1.0, q_seed=seeds, use_standoff=False, standoff_offset=-1, w_vel=0.0,
                               field_all=b.field_obs.copy(), flags=capi.FLAG_OBS_LINEAR if linear_obstacle else 0)


@pytest.mark.parametrize("linear_obstacle", [False, True])
def test_ik_batch_matches_oracle(linear_obstacle):
    """SURVEY 8(f) row 1 against the oracle: batched IK of 48 Panda grasps (7 joints, 6-DoF goal: one redundant direction, so the
    joint vector is not unique but the gripper point positions and the objective are) -- CUDA vs the C oracle on identical inputs;
    ``linear_obstacle``: the reference IK solver's unsquared collision term 10*sum(c) (gto/ik_solver.py:69)."""
    import c_oracle
    import gto_oracle as O
    from helpers import small_workload, upload_fields, problems_from_workload
    from grasptrajopt_b200 import kinematics as K

    w = small_workload("C2", "panda_small", B=48, n_field=64)
    b = _ik_batch_from_workload(w, 48, linear_obstacle)
    w.batch = b
    ctx = capi.GtoContext(0)
    try:
        ctx.set_robot(w.table)
        upload_fields(ctx, w)
        opts = capi.default_options(max_iter=50)  # reference: max_iter 50 (:75)
        res = ctx.solve_batch(b, opts)
    finally:
        ctx.close()
    ora = c_oracle.solve_workload(w, options=c_oracle.default_options(max_iter=50))
    t = w.table
    xg = t.points[t.grip_pt_start : t.grip_pt_start + t.grip_pt_count]

    def gripper_points(q):
        F = np.stack([O.gripper_frame(t, qi) for qi in q])
        return np.einsum("bij,kj->bki", F[:, :3, :3], xg) + F[:, None, :3, 3]

    Wg, Wo = gripper_points(res["Q"][:, 2]), gripper_points(ora["Q"][:, 2])
    both = (res["status"] == 0) & (ora["status"] == 0)
    # (the unsquared term has a piecewise-constant gradient: most of its solutions rest on a cell face and are not "converged")
    assert both.sum() >= (8 if linear_obstacle else 24), (res["status"], ora["status"])
    assert np.abs(Wg - Wo)[both].max() < 1e-4  # metres
    rel = np.abs(res["cost"] - ora["cost"]) / np.maximum(ora["cost"], 1e-9)
    assert rel[both].max() < 1e-4
    assert np.mean((res["status"] == 0) == (ora["status"] == 0)) >= 0.9
    # the NumPy oracle (the one pinned against the reference's own code) agrees with the C port on this problem family
    p0 = problems_from_workload(w, [0, 5])
    for i, p in zip((0, 5), p0):
        r = O.solve_lm(p, O.SolverOptions(max_iter=50))
        assert r.status == ora["status"][i] and abs(r.cost - ora["cost"][i]) <= 1e-9 * max(1.0, r.cost)
    if linear_obstacle:  # objective = goal term + 10 * sum(c) at the solution (trilinear c), not 10 * sum(c^2)
        p = p0[0]
        lin = O.linearize(p, ora["Q"][0], need_jac=False)
        goal = float(np.sum(lin.r_goal ** 2))
        c = lin.r_obs[2] / np.sqrt(p.w_obs)
        # knots 0,1 (pinned at the seed) carry their constant obstacle cost as well
        c_all = lin.r_obs / np.sqrt(p.w_obs)
        assert ora["cost"][0] == pytest.approx(goal + p.w_obs * float(c_all.sum()), rel=1e-9)
        assert c.sum() >= 0
