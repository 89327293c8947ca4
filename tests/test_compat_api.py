"""Host-side mirror of the reference interface (optas / gto / mesh_to_sdf packages).  CPU only."""
import os
import struct

import numpy as np
import pytest

import optas
from gto.gto_models import GTORobotModel
from gto.utils import interpolate_waypoints
from mesh_to_sdf.depth_point_cloud import DepthPointCloud
from grasptrajopt_b200.robot_table import RobotTable
import gto_oracle as O

from conftest import GOLDEN, ASSETS

URDF = """<?xml version="1.0"?>
<robot name="arm3">
  <link name="base"><visual><geometry><mesh filename="box.obj"/></geometry></visual></link>
  <link name="l1"><visual><origin xyz="0 0 0.1" rpy="0 0 0.3"/><geometry><mesh filename="box.obj"/></geometry></visual></link>
  <link name="l2"><visual><geometry><mesh filename="box.stl"/></geometry></visual></link>
  <link name="tool"><visual><geometry><mesh filename="box.obj" scale="0.5 0.5 0.5"/></geometry></visual></link>
  <link name="finger"/>
  <joint name="j1" type="revolute"><parent link="base"/><child link="l1"/><origin xyz="0 0 0.2" rpy="0 0 0"/><axis xyz="0 0 1"/>
    <limit lower="-2" upper="2" velocity="1" effort="1"/></joint>
  <joint name="j2" type="continuous"><parent link="l1"/><child link="l2"/><origin xyz="0.3 0 0" rpy="1.2 0 0"/><axis xyz="0 1 0"/></joint>
  <joint name="fix" type="fixed"><parent link="l2"/><child link="tool"/><origin xyz="0.25 0 0" rpy="0 0.4 0"/></joint>
  <joint name="slide" type="prismatic"><parent link="tool"/><child link="finger"/><origin xyz="0 0 0.05"/><axis xyz="0 2 0"/>
    <limit lower="0" upper="0.04" velocity="1" effort="1"/></joint>
</robot>
"""


def _write_box(tmp_path):
    v = np.array([[x, y, z] for x in (-0.05, 0.05) for y in (-0.04, 0.04) for z in (-0.1, 0.1)])
    f = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    with open(tmp_path / "box.obj", "w") as fh:
        for p in v:
            fh.write(f"v {p[0]} {p[1]} {p[2]}\n")
        for q in f:
            fh.write("f " + " ".join(f"{i + 1}/{i + 1}" for i in q) + "\n")
    tris = []
    for q in f:
        tris += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
    with open(tmp_path / "box.stl", "wb") as fh:
        fh.write(b"\0" * 80 + struct.pack("<I", len(tris)))
        for t in tris:
            fh.write(struct.pack("<3f", 0, 0, 0))
            for i in t:
                fh.write(struct.pack("<3f", *v[i]))
            fh.write(b"\0\0")
    (tmp_path / "arm3.urdf").write_text(URDF)


@pytest.fixture()
def model(tmp_path):
    _write_box(tmp_path)
    return GTORobotModel(str(tmp_path), urdf_filename=str(tmp_path / "arm3.urdf"), time_derivs=[0, 1], param_joints=["slide"],
                         collision_link_names=["base", "l1", "l2", "tool"], sample_point_count=40, seed=3)


def test_robot_model_properties(model):
    assert model.ndof == 3 and model.get_name() == "arm3"
    assert model.actuated_joint_names == ["j1", "j2", "slide"]
    assert model.optimized_joint_names == ["j1", "j2"] and model.parameter_joint_indexes == [2]
    np.testing.assert_allclose(model.lower_actuated_joint_limits.toarray().ravel(), [-2, -1e9, 0])
    np.testing.assert_allclose(model.upper_optimized_joint_limits.toarray().ravel(), [2, 1e9])
    Q = optas.DM(np.arange(6.0).reshape(3, 2))
    assert model.extract_optimized_dimensions(Q).toarray().shape == (2, 2)
    assert model.extract_parameter_dimensions(Q).toarray().tolist() == [[4.0, 5.0]]
    assert model.get_root_link() == "base"


def test_fk_functions_and_table_agree(model):
    """Three implementations of the same chain: RobotModel FK (URDF walk), visual_tf, and the flattened table (oracle)."""
    t = model.to_table("tool", "tool")
    assert t.nopt == 2 and t.nmov == 2 and t.nlinks == 4  # the parameter joint below the tool is not an ancestor of any link
    q = np.array([0.7, -1.1, 0.02])
    F = O.link_frames(t, q)
    for l, name in enumerate(t.link_names):
        np.testing.assert_allclose(F[l], model.visual_tf[name](q).toarray(), atol=1e-12)
    np.testing.assert_allclose(O.gripper_frame(t, q), model.get_global_link_transform("tool", q).toarray(), atol=1e-12)
    fn = model.get_global_link_transform_function("tool", n=2)
    out = fn(np.stack([q, q + 0.1], axis=1))
    assert isinstance(out, list) and len(out) == 2
    np.testing.assert_allclose(out[0].toarray(), O.gripper_frame(t, q), atol=1e-12)
    pts, nrm = model.compute_fk_surface_points(q)
    np.testing.assert_allclose(pts, O.world_points(t, q), atol=1e-12)
    assert pts.shape == (160, 3) and np.allclose(np.linalg.norm(nrm, axis=1), 1.0)
    # geometric Jacobian vs the twist form the kernels use
    J = model.get_global_link_geometric_jacobian("tool", q).toarray()
    om, mm = O.joint_twists(t, O.fk_movable(t, q))
    e = O.gripper_frame(t, q)[:3, 3]
    for k in range(2):
        np.testing.assert_allclose(J[:3, k], np.cross(om[k], e) + mm[k], atol=1e-12)


def test_surface_sampling_is_seeded_and_on_the_surface(model, tmp_path):
    again = GTORobotModel(str(tmp_path), urdf_filename=str(tmp_path / "arm3.urdf"), param_joints=["slide"],
                          collision_link_names=["base", "l1", "l2", "tool"], sample_point_count=40, seed=3)
    for name in model.surface_pc_map:
        np.testing.assert_array_equal(model.surface_pc_map[name].points, again.surface_pc_map[name].points)
    p = model.surface_pc_map["l1"].points
    on_face = np.isclose(np.abs(p[:, 0]), 0.05) | np.isclose(np.abs(p[:, 1]), 0.04) | np.isclose(np.abs(p[:, 2]), 0.1)
    assert on_face.all()
    assert np.abs(model.surface_pc_map["tool"].points).max() <= 0.05 + 1e-12  # mesh scale 0.5 applied


def test_field_geometry_matches_reference_run(model):
    z = np.load(os.path.join(GOLDEN, "ref_field.npz"))
    model.setup_workspace_field(arm_len=1.0, arm_height=0)
    np.testing.assert_allclose(model.origin, z["ws_origin"])
    assert tuple(model.field_shape) == tuple(z["ws_shape"]) and model.field_size == int(z["ws_size"])
    np.testing.assert_array_equal(model.points_to_offsets_numpy(z["query"].copy()), z["offsets_numpy"])
    model.setup_points_field(np.stack([z["pf_cloud_min"], z["pf_cloud_max"]]))
    np.testing.assert_allclose(model.origin, z["pf_origin"])
    assert tuple(model.field_shape) == tuple(z["pf_shape"])


def test_depth_point_cloud_oracle_matches_reference_run():
    """oracle/dpc_oracle.py (the checker of the GPU DepthPointCloud) against the outputs of the reference's own class."""
    from dpc_oracle import KDTreeDepthPointCloud

    z = np.load(os.path.join(GOLDEN, "ref_field.npz"))
    dpc = KDTreeDepthPointCloud(z["dpc_depth"], z["dpc_K"], z["dpc_cam"], target_mask=None, threshold=1.5)
    np.testing.assert_allclose(dpc.points, z["dpc_points"], atol=1e-12)
    np.testing.assert_array_equal(dpc.get_sdf(z["dpc_query"]), z["dpc_sdf"])
    np.testing.assert_array_equal(dpc.get_sdf_cost(z["dpc_query"], epsilon=0.02), z["dpc_cost"])


def test_interpolate_waypoints_matches_reference_run():
    z = np.load(os.path.join(GOLDEN, "ref_seed.npz"))
    np.testing.assert_allclose(interpolate_waypoints(np.stack([z["qc"], z["qg"]]), 30, 9), z["cubic_T30"], atol=1e-14)
    np.testing.assert_allclose(interpolate_waypoints(np.stack([z["qc"], z["qg"]]), 50, 9, mode="linear"), z["linear_T50"], atol=1e-14)
    np.testing.assert_allclose(interpolate_waypoints(np.stack([z["qc"], z["mid"], z["qg"]]), 30, 9), z["cubic3_T30"], atol=1e-14)


def test_dm_stand_in_behaves_like_casadi_for_the_callers():
    qc = [0.1, 0.2, 0.3]
    Q0 = optas.diag(qc) @ optas.DM.ones(3, 5)  # gto_planner.py:152
    assert Q0.shape == (3, 5) and np.allclose(Q0.toarray()[:, 3], qc)
    Q0[:, 4] = np.array([1.0, 2.0, 3.0])  # gto_planner.py:219
    assert Q0.toarray()[:, 4].tolist() == [1.0, 2.0, 3.0]
    t = optas.linspace(0, 10.0, 50)  # gto_planner.py:27-28
    assert float((t[1] - t[0]).toarray()[0, 0]) == pytest.approx(10.0 / 49)
    assert optas.DM([1, 2, 3]).shape == (3, 1)
    with pytest.raises(AttributeError):
        optas.CasADiSolver


def test_planner_keeps_reference_signature():
    import inspect
    from gto.gto_planner import GTOPlanner
    from gto.ik_solver import IKSolver

    assert list(inspect.signature(GTOPlanner.__init__).parameters)[:7] == ["self", "robot", "link_ee", "link_gripper", "collision_avoidance", "standoff_distance", "standoff_offset"]
    assert list(inspect.signature(GTOPlanner.plan).parameters) == ["self", "qc", "RT", "sdf_cost_obstacle", "base_position", "q_solution", "use_standoff", "axis_standoff"]
    assert list(inspect.signature(GTOPlanner.plan_goalset).parameters) == ["self", "qc", "RTs", "sdf_cost_all", "sdf_cost_obstacle", "base_position", "q_solutions", "use_standoff", "axis_standoff", "interpolate"]
    assert list(inspect.signature(GTOPlanner.setup_optimization).parameters) == ["self", "goal_size", "use_standoff", "axis_standoff"]
    assert list(inspect.signature(IKSolver.solve_ik).parameters) == ["self", "q_0", "RT", "sdf_cost_obstacle", "base_position"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/data/robots/panda"), reason="reference checkout not present")
def test_reference_urdfs_load_and_match_assets():
    from gto.utils import load_yaml

    cfg = load_yaml("/root/reference/data/configs/panda.yaml")["robot_cfg"]
    robot = GTORobotModel("/root/reference/data/robots/panda", urdf_filename="/root/reference/data/robots/panda/panda.urdf",
                          time_derivs=[0, 1], param_joints=cfg["param_joints"], collision_link_names=cfg["collision_link_names"],
                          sample_point_count=8, seed=0)
    t = robot.to_table(cfg["link_ee"], cfg["link_gripper"])
    ref = RobotTable.load(os.path.join(ASSETS, "panda_c2.npz"))
    assert t.mov_names == ref.mov_names and t.link_names == ref.link_names
    np.testing.assert_allclose(t.mov_origin, ref.mov_origin)
    np.testing.assert_allclose(t.lo, ref.lo)
    assert robot.ndof == 9 and robot.optimized_joint_names[0] == "panda_joint1"


def test_base_planner_surface(model):
    """gto.BasePlanner keeps the reference's constructor / setup_optimization / plan_goalset names (gto/base_planner.py:19-94);
    without a GPU the call fails loudly instead of falling back to anything."""
    from gto.base_planner import BasePlanner
    from grasptrajopt_b200 import capi

    bp = BasePlanner(model, "tool", "tool")
    assert bp.task_name == "base_pose_estimator" and bp.gripper_points.shape[1] == 3
    bp.setup_optimization(goal_size=2, base_effort_weight=0.02)
    assert bp.goal_size == 2 and bp.base_effort_weight == 0.02
    with pytest.raises(ValueError):
        bp.setup_optimization(goal_size=40)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises((capi.GtoError, capi.GtoLibraryError)):
            bp.plan_goalset(np.zeros(3), np.tile(np.eye(4), (2, 1, 1)))


def test_base_planner_pose_errors(model):
    """BasePlanner._errors (reference gto/base_planner.py:131-148): zero for goals that are exactly the gripper poses seen from the new
    base, the applied offsets otherwise."""
    from gto.base_planner import BasePlanner
    from gto.utils import rotZ

    bp = BasePlanner(model, "tool", "tool")
    bp.setup_optimization(goal_size=2)
    y = np.array([[0.3, -0.2, 0.4]])
    Tb = rotZ(y[0, 2]); Tb[0, 3], Tb[1, 3] = y[0, 0], y[0, 1]
    Q = np.array([[[0.5, 0.3, 0.01], [-0.7, 1.1, 0.01]]])
    RTs = np.stack([np.linalg.inv(Tb) @ model.get_global_link_transform("tool", q).toarray() for q in Q[0]])[None]
    ep, er = bp._errors(Q, y, RTs)
    assert ep.shape == (1, 2) and np.abs(ep).max() < 1e-6 and np.abs(er).max() < 1e-2
    RTs2 = RTs.copy()
    RTs2[0, 0, :3, 3] += np.linalg.inv(Tb)[:3, :3] @ np.array([0.0, 0.0, 0.05])  # 5 cm along the new base's z
    RTs2[0, 1, :3, :3] = RTs2[0, 1, :3, :3] @ rotZ(np.radians(10.0))[:3, :3]
    ep2, er2 = bp._errors(Q, y, RTs2)
    assert ep2[0, 0] == pytest.approx(0.05, abs=1e-6) and er2[0, 1] == pytest.approx(10.0, abs=1e-2)


def test_plan_collision_audit_host_logic(model):
    """gto.utils.plan_collision_audit against the reference's per-knot loop (examples/pybullet_evaluate_plans.py:219-237); the
    distance queries come from the KD-tree oracle here (the GPU class is exercised in tests/test_gpu_cloud.py)."""
    from gto.utils import plan_collision_audit
    from dpc_oracle import KDTreeDepthPointCloud

    H, Wd, f = 60, 80, 70.0
    K = np.array([[f, 0, Wd / 2], [0, f, H / 2], [0, 0, 1.0]])
    cam = np.eye(4); cam[:3, :3] = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]]); cam[:3, 3] = [0.3, 0.0, 1.0]
    depth = np.full((H, Wd), 1.0, np.float32)
    depth[20:40, 30:60] = 0.8  # a box on the table
    dpc = KDTreeDepthPointCloud(depth, K, cam, threshold=1.5)
    T = 6
    plan = np.stack([np.linspace(-1.5, 1.5, T), np.linspace(0.0, 3.0, T), np.full(T, 0.01)])
    base = np.array([0.2, 0.0, 0.12])
    hit, first, counts = plan_collision_audit(model, plan, dpc, base, min_points=5)
    ref = np.array([(dpc.get_sdf(model.compute_fk_surface_points(plan[:, i])[0] + base) < 0).sum() for i in range(T)])
    assert np.array_equal(counts, ref)
    assert hit == bool((ref > 5).any()) and first == (int(np.flatnonzero(ref > 5)[0]) if (ref > 5).any() else -1)
    hit0, first0, c0 = plan_collision_audit(model, plan, dpc, np.array([0.2, 0.0, 5.0]))  # far above everything
    assert not hit0 and first0 == -1 and c0.sum() == 0
