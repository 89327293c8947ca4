"""The acceptance driver of the north star: the reference's ``examples/pybullet_gto_planning.py`` runs UNCHANGED (no edit, no copy)
through ``python -m grasptrajopt_b200.run_reference``.  The simulator side (PyBullet, SceneReplica data, matplotlib,
transforms3d -- all absent here) is replaced by stub modules; the GPU-side classes are observed at the call boundary (this is a
CPU test: the captured ``plan_goalset`` call is the one the script makes, with the arguments it built from the compat
``GTORobotModel`` / ``IKSolver`` / ``DepthPointCloud``).  Skipped where the reference checkout is not present (GPU box)."""
import os
import sys
import types

import numpy as np
import pytest

REF = os.environ.get("GTO_REFERENCE_DIR", "/root/reference")
SCRIPT = os.path.join(REF, "examples", "pybullet_gto_planning.py")
pytestmark = pytest.mark.skipif(not os.path.exists(SCRIPT), reason="reference checkout not present")


def _rot_to_quat_wxyz(R):
    from grasptrajopt_b200.spatial import mat2quat_wxyz

    return mat2quat_wxyz(np.asarray(R))


def _quat_wxyz_to_rot(q):
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class _FakeRobot:
    """What the script uses of ``pybullet_api.Panda`` (examples/pybullet_api.py:223-260,340)."""

    ndof = 9

    def __init__(self, log):
        self.log = log

    def q(self):
        return [0.0, -1.285, 0.0, -2.356, 0.0, 1.571, 0.785, 0.04, 0.04]

    def get_standoff_pose(self, offset, axis):
        T = np.eye(4, dtype=np.float32)
        T["xyz".index(axis), 3] = offset
        return T

    def execute_plan(self, plan):
        self.log.append(("execute_plan", np.asarray(plan).shape))

    def close_gripper(self):
        self.log.append(("close_gripper",))

    def retract(self):
        self.log.append(("retract",))


class _FakeEnv:
    """What the script uses of ``pybullet_scenereplica.SceneReplicaEnv``: one scene, one object on a table seen from above."""

    def __init__(self, urdf_filename, data_dir, robot_name, scene_type, log):
        self.log = log
        self.robot = _FakeRobot(log)
        self.all_scene_ids = [10]
        self.base_position = np.array([0.0, 0.0, 0.0])
        self.object_names, self.object_uids = ["003_cracker_box"], [7]

    def setup_scene(self, scene_id):
        return {"nearest_first": ["003_cracker_box"], "random": ["003_cracker_box"]}

    def reset_scene(self, names):
        pass

    def get_observation(self):
        H, W, f = 48, 64, 60.0
        K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
        cam = np.eye(4)
        cam[:3, :3] = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]])  # looking straight down
        cam[:3, 3] = [0.5, 0.0, 1.0]
        depth = np.full((H, W), 1.0, np.float32)  # table at z = 0
        mask = np.zeros((H, W), np.int32)
        depth[20:28, 28:36] = 0.85  # the object: a box of 15 cm height
        mask[20:28, 28:36] = 7
        return np.zeros((H, W, 4), np.uint8), depth, mask, cam, K

    def get_object_pose(self, name):
        return (0.5, 0.0, 0.075), (0.0, 0.0, 0.0, 1.0)  # position, quaternion xyzw

    def record_gripper_position(self):
        pass

    def retract(self, distance):
        self.log.append(("env.retract", distance))

    def compute_reward(self, name):
        return 1

    def reset_objects(self, name):
        pass


def test_reference_script_runs_unchanged(tmp_path, monkeypatch):
    import grasptrajopt_b200
    from grasptrajopt_b200 import run_reference

    log, captured = [], {}
    # ---- stub modules for the simulator side ----
    stubs = {}
    stubs["pybullet"] = types.ModuleType("pybullet")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    stubs["matplotlib"], stubs["matplotlib.pyplot"] = mpl, mpl.pyplot
    t3d = types.ModuleType("transforms3d")
    t3d.quaternions = types.ModuleType("transforms3d.quaternions")
    t3d.quaternions.mat2quat, t3d.quaternions.quat2mat = _rot_to_quat_wxyz, _quat_wxyz_to_rot
    t3d.euler = types.ModuleType("transforms3d.euler")
    t3d.euler.mat2euler = lambda R: (0.0, 0.0, 0.0)
    stubs["transforms3d"], stubs["transforms3d.quaternions"], stubs["transforms3d.euler"] = t3d, t3d.quaternions, t3d.euler
    api = types.ModuleType("pybullet_api")
    api.Fetch, api.Panda = _FakeRobot, _FakeRobot
    stubs["pybullet_api"] = api
    sr = types.ModuleType("pybullet_scenereplica")
    sr.SceneReplicaEnv = lambda urdf, data_dir, robot_name, scene_type: _FakeEnv(urdf, data_dir, robot_name, scene_type, log)
    stubs["pybullet_scenereplica"] = sr
    for name, mod in stubs.items():
        monkeypatch.setitem(sys.modules, name, mod)
    for name in ("utils", "_init_paths"):  # the script's own helper modules must be imported fresh from the reference's examples/
        monkeypatch.delitem(sys.modules, name, raising=False)
    monkeypatch.setattr(sys, "path", list(sys.path))
    monkeypatch.setenv("GTO_ROOT_DIR", REF)

    # ---- GPU-side classes observed at the call boundary (CPU test) ----
    grasptrajopt_b200.install_compat()
    import gto.gto_planner as GP
    import gto.ik_solver as IK
    import mesh_to_sdf.depth_point_cloud as DPC
    from dpc_oracle import KDTreeDepthPointCloud

    class _CloudOnCPU(KDTreeDepthPointCloud):  # same constructor / methods as the product class, distances from the KD-tree oracle
        def __init__(self, depth, intrinsic_matrix, camera_pose, target_mask=None, threshold=1.5):
            super().__init__(depth, intrinsic_matrix, camera_pose, target_mask, threshold)
            captured.setdefault("clouds", []).append(self.points.shape)

    monkeypatch.setattr(DPC, "DepthPointCloud", _CloudOnCPU)

    def fake_solve_ik(self, q_0, RT, sdf_cost_obstacle, base_position):
        captured.setdefault("ik_calls", []).append(np.asarray(RT).copy())
        assert self.solver is not None  # setup_optimization() was called by the script
        return np.asarray(q_0, dtype=np.float64).reshape(-1), 0.0, 0.0, 0.0

    def fake_plan_goalset(self, qc, RTs, sdf_cost_all, sdf_cost_obstacle, base_position, q_solutions=None, use_standoff=True, axis_standoff="x",
                          interpolate=True):
        captured["plan_goalset"] = dict(qc=np.asarray(qc), RTs=np.asarray(RTs), sdf_cost_all=np.asarray(sdf_cost_all), sdf_cost_obstacle=np.asarray(sdf_cost_obstacle),
                                        base_position=np.asarray(base_position), q_solutions=np.asarray(q_solutions), use_standoff=use_standoff,
                                        axis_standoff=axis_standoff, interpolate=interpolate, T=self.T, field_size=self.robot.field_size,
                                        standoff=(self.standoff_distance, self.standoff_offset))
        Q = np.tile(np.asarray(qc, dtype=np.float64).reshape(-1, 1), (1, self.T))
        return Q, np.zeros((Q.shape[0], self.T - 1)), np.array([0.0])

    monkeypatch.setattr(IK.IKSolver, "solve_ik", fake_solve_ik)
    monkeypatch.setattr(GP.GTOPlanner, "plan_goalset", fake_plan_goalset)
    import time as _time
    monkeypatch.setattr(_time, "sleep", lambda s: None)

    # ---- data the script loads: two grasps for the object (examples/pybullet_gto_planning.py:21-45) ----
    data_dir = tmp_path / "data"
    gdir = data_dir / "grasp_data" / "panda_simulated"
    gdir.mkdir(parents=True)
    g = np.tile(np.eye(4), (2, 1, 1))
    g[:, :3, :3] = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]])  # gripper pointing down
    g[0, :3, 3] = [0.0, 0.0, 0.25]
    g[1, :3, 3] = [0.02, 0.0, 0.3]
    np.save(gdir / "003_cracker_box.npy", {"transforms": g}, allow_pickle=True)
    monkeypatch.chdir(tmp_path)

    # ---- run the unmodified script through the launcher ----
    with open(SCRIPT, "rb") as fh:
        before = fh.read()
    run_reference.main([SCRIPT, "--robot", "panda", "--scene_type", "tabletop", "--scene_id", "10", "-d", str(data_dir)])
    with open(SCRIPT, "rb") as fh:
        assert fh.read() == before  # nothing in the reference tree was touched

    # the packages the script imported are the CasADi-free ones of this repository, not the reference's
    import gto, optas, mesh_to_sdf
    for m in (gto, optas, mesh_to_sdf):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(grasptrajopt_b200.COMPAT_DIR)), m.__file__
    assert os.path.join(os.path.dirname(SCRIPT), "..") in sys.path  # ... although the checkout root is on the path (for data/)

    # the captured call: both orderings plan once for the object, with the reference's argument conventions
    c = captured["plan_goalset"]
    assert c["qc"].shape == (9,) and c["RTs"].ndim == 3 and c["RTs"].shape[1:] == (4, 4) and 1 <= c["RTs"].shape[0] <= 2
    assert c["q_solutions"].shape == (9, c["RTs"].shape[0]) and c["q_solutions"].dtype == np.float32  # Q12: float32 IK seeds
    assert c["sdf_cost_all"].shape == (c["field_size"],) == c["sdf_cost_obstacle"].shape
    assert c["sdf_cost_all"].max() > 0 and c["sdf_cost_obstacle"].max() > 0
    assert (c["sdf_cost_all"] >= c["sdf_cost_obstacle"] - 1e-6).mean() > 0.99  # the target is masked out of the obstacle field
    assert c["use_standoff"] is True and c["axis_standoff"] == "z" and c["interpolate"] is True and c["T"] == 50
    assert c["standoff"] == (-0.1, -10)
    assert len(captured["ik_calls"]) >= 2 and len(captured["clouds"]) == 4  # 2 orderings x (all, obstacle) clouds
    assert ("execute_plan", (9, 50)) in log and ("close_gripper",) in log and ("env.retract", 0.3) in log or any(e[0] == "env.retract" for e in log)
    out = list((tmp_path / "results").glob("GTO_scenereplica_panda_tabletop_*.json"))
    assert len(out) == 1
    import json
    res = json.load(open(out[0]))
    assert res["10"]["nearest_first"]["003_cracker_box"]["reward"] == 1 and len(res["10"]["random"]["003_cracker_box"]["plan"]) == 9
