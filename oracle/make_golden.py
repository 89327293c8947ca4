"""Generate ``tests/golden/ref_*.npz`` by executing the reference's own Python.

TEST INFRASTRUCTURE.  Run in the authoring container only (needs ``/root/reference``):

    python oracle/make_golden.py [/root/reference]

The reference cannot be imported as a package (``optas/__init__.py`` needs CasADi/VTK/...).  Its
modules on the hot path are therefore loaded one by one from where they lie, with numeric stand-ins
registered for the third-party imports they make (``oracle/refshim``): ``casadi`` -> ndarray-backed DM,
``urdf_parser_py`` -> this repo's URDF reader, ``xacro``/``trimesh``/``pyrender``/``transforms3d``/``turtle``
-> empty stubs.  What is executed from the reference and stored:

  ref_fk.npz      optas/spatialmath.py + optas/models.py  RobotModel.get_global_link_transform,
                  get_link_transform, get_link_visual_origin, joint-limit/index properties   (A1, A2, A11)
  ref_sdf.npz     gto/sdf_callback.py  SDFCallback / JacFun / HesFun .eval                        (A5, A6)
  ref_field.npz   gto/gto_models.py  setup_workspace_field, setup_points_field,
                  points_to_offsets_numpy, compute_plan_cost; mesh_to_sdf/depth_point_cloud.py
                  DepthPointCloud.get_sdf / get_sdf_cost                                         (A5, A7, A14)
  ref_seed.npz    gto/utils.py  interpolate_waypoints                                            (A14)
  ref_stored_plans.npz   a sample of the plans stored in examples/results_iros2024/*.json         (section 4)

No reference source is copied; only inputs and outputs are written.
"""
import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, "refshim"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference(ref):
    import casadi_numeric
    from grasptrajopt_b200 import urdf as my_urdf

    sys.modules["casadi"] = casadi_numeric
    _stub("xacro")
    _stub("urdf_parser_py")
    _stub("urdf_parser_py.urdf", URDF=my_urdf.URDF, Joint=my_urdf.Joint, Link=my_urdf.Link, Pose=my_urdf.Pose)
    _stub("trimesh")
    _stub("pyrender")
    _stub("turtle", color=None)
    _stub("_init_paths")
    _stub("transforms3d")
    _stub("transforms3d.quaternions", quat2mat=None, mat2quat=None)
    for pkg in ("optas", "gto", "mesh_to_sdf"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(ref, pkg)]
        sys.modules[pkg] = m
    _stub("optas.visualize", Visualizer=None)
    mods = {}
    for name in ("optas.spatialmath", "optas.models", "gto.utils", "gto.sdf_callback", "mesh_to_sdf.depth_point_cloud"):
        mods[name] = importlib.import_module(name)
    # gto.gto_models does ``import optas`` / ``import mesh_to_sdf`` and uses optas.floor only in the symbolic path
    mods["gto.gto_models"] = importlib.import_module("gto.gto_models")
    return mods


def gen_fk(ref, mods, rng):
    import yaml

    RobotModel = mods["optas.models"].RobotModel
    sm = mods["optas.spatialmath"]
    out = {}
    for robot in ("panda", "fetch"):
        cfg = yaml.safe_load(open(os.path.join(ref, "data", "configs", f"{robot}.yaml")))["robot_cfg"]
        model = RobotModel(urdf_filename=os.path.join(ref, cfg["urdf_robot_path"]), time_derivs=[0, 1], param_joints=cfg["param_joints"])
        lo = model.lower_actuated_joint_limits.toarray().reshape(-1)
        hi = model.upper_actuated_joint_limits.toarray().reshape(-1)
        K = 6
        q = np.zeros((K, model.ndof))
        q[0] = np.array(cfg["default_pose"])
        for k in range(1, K):
            q[k] = np.where(hi > lo, rng.uniform(np.maximum(lo, -3.0), np.minimum(hi, 3.0)), rng.uniform(-1, 1, size=lo.shape))
        links = list(cfg["collision_link_names"])
        TL = np.zeros((K, len(links), 4, 4))
        TV = np.zeros((K, len(links), 4, 4))
        G = np.zeros((K, 4, 4))
        TE = np.zeros((K, 4, 4))
        for k in range(K):
            for i, name in enumerate(links):
                lnk = model.get_global_link_transform(name, q[k])
                xyz, rpy = model.get_link_visual_origin(model.urdf.link_map[name])
                vis = sm.rt2tr(sm.rpy2r(rpy), xyz)  # gto/gto_models.py:96-100
                TL[k, i] = lnk.toarray()
                TV[k, i] = (lnk @ vis).toarray()
            G[k] = model.get_link_transform(cfg["link_gripper"], q[k], cfg["link_ee"]).toarray()
            TE[k] = model.get_global_link_transform(cfg["link_ee"], q[k]).toarray()
        out.update(
            {
                f"{robot}_q": q,
                f"{robot}_links": np.array(links),
                f"{robot}_T_link": TL,
                f"{robot}_T_visual": TV,
                f"{robot}_G": G,
                f"{robot}_T_ee": TE,
                f"{robot}_actuated": np.array(model.actuated_joint_names),
                f"{robot}_opt_idx": np.array(model.optimized_joint_indexes),
                f"{robot}_par_idx": np.array(model.parameter_joint_indexes),
                f"{robot}_lo": lo,
                f"{robot}_hi": hi,
                f"{robot}_lo_opt": model.lower_optimized_joint_limits.toarray().reshape(-1),
                f"{robot}_hi_opt": model.upper_optimized_joint_limits.toarray().reshape(-1),
            }
        )
    # spatialmath primitives
    rpy = rng.uniform(-3, 3, size=(5, 3))
    out["sm_rpy"] = rpy
    out["sm_rpy2r"] = np.stack([sm.rpy2r(r).toarray() for r in rpy])
    ang = rng.uniform(-3, 3, size=5)
    axs = rng.normal(size=(5, 3))
    out["sm_ang"], out["sm_axis"] = ang, axs
    out["sm_angvec2r"] = np.stack([sm.angvec2r(a, v).toarray() for a, v in zip(ang, axs)])
    out["sm_standoff_z"] = sm.standoff(-0.1, "z").toarray()
    out["sm_standoff_x"] = sm.standoff(-0.2, "x").toarray()
    Tr = sm.rt2tr(sm.rpy2r(rpy[0]), [0.1, -0.2, 0.3])
    out["sm_T"] = Tr.toarray()
    out["sm_invt"] = sm.invt(Tr).toarray()
    np.savez_compressed(os.path.join(OUT, "ref_fk.npz"), **out)
    print("ref_fk.npz", {k: v.shape for k, v in out.items() if k.endswith("T_link")})


def gen_sdf(mods, rng):
    cb = mods["gto.sdf_callback"]
    origin = np.array([-0.2, -1.3, -0.2])  # the reference's own demo geometry (sdf_callback.py:189-199)
    shape = (15, 26, 15)
    res = 0.1
    data = rng.random(shape).reshape(-1)
    f = cb.SDFCallback("f", data, origin, res, shape)
    jac = f.get_jacobian("jac_f", None, None, {})
    hes = jac.get_jacobian("hes_f", None, None, {})
    lo = origin - 0.3
    hi = origin + res * np.array(shape) + 0.3
    pts = rng.uniform(lo, hi, size=(96, 3))
    pts[:8] = origin + res * rng.integers(0, 10, size=(8, 3))  # exactly on nodes
    val = np.array([float(np.asarray(f.eval([p.reshape(3, 1)])[0]).reshape(-1)[0]) for p in pts])
    J = np.stack([np.asarray(jac.eval([p.reshape(3, 1), None])[0]).reshape(3) for p in pts])
    H = np.stack([np.asarray(hes.eval([p.reshape(3, 1), None, None])[0]).reshape(3, 3) for p in pts])
    np.savez_compressed(os.path.join(OUT, "ref_sdf.npz"), origin=origin, shape=np.array(shape), pitch=res, data=data, points=pts, value=val, jac=J, hess=H)
    print("ref_sdf.npz", val.shape, J.shape, H.shape)


def gen_field(ref, mods, rng):
    from types import SimpleNamespace
    from grasptrajopt_b200.robot_table import RobotTable

    gm = mods["gto.gto_models"]
    RobotModel = mods["optas.models"].RobotModel
    sm = mods["optas.spatialmath"]
    import yaml

    cfg = yaml.safe_load(open(os.path.join(ref, "data", "configs", "panda.yaml")))["robot_cfg"]
    table = RobotTable.load(os.path.join(REPO, "grasptrajopt_b200", "assets", "panda_small.npz"))
    robot = object.__new__(gm.GTORobotModel)  # skip trimesh/CasADi-symbolic construction
    RobotModel.__init__(robot, urdf_filename=os.path.join(ref, cfg["urdf_robot_path"]), time_derivs=[0, 1], param_joints=cfg["param_joints"])
    robot.collision_link_names = cfg["collision_link_names"]
    robot.field_margin, robot.grid_resolution = 0.4, 0.05
    robot.surface_pc_map = {}
    robot.visual_tf = {}
    for l, name in enumerate(table.link_names):
        s, c = int(table.link_pt_start[l]), int(table.link_pt_count[l])
        robot.surface_pc_map[name] = SimpleNamespace(points=table.points[s : s + c], normals=np.zeros((c, 3)))
        xyz, rpy = robot.get_link_visual_origin(robot.urdf.link_map[name])
        vis = sm.rt2tr(sm.rpy2r(rpy), xyz)
        robot.visual_tf[name] = (lambda nm, v: (lambda q: robot.get_global_link_transform(nm, q) @ v))(name, vis)
    out = {}
    robot.setup_workspace_field(arm_len=cfg["arm_len"], arm_height=cfg["arm_height"])
    out["ws_origin"], out["ws_shape"], out["ws_size"] = robot.origin.copy(), np.array(robot.field_shape), robot.field_size
    out["ws_points_first"], out["ws_points_last"] = robot.workspace_points[0], robot.workspace_points[-1]
    pts = rng.uniform([-0.6, -1.6, -0.6], [1.6, 1.6, 1.6], size=(200, 3))
    out["query"] = pts
    out["offsets_numpy"] = robot.points_to_offsets_numpy(pts.copy())
    field = rng.random(robot.field_size).astype(np.float32)
    out["field"] = field
    qc = np.array(cfg["default_pose"])
    qg = qc.copy()
    qg[:7] += np.array([0.4, 0.5, -0.3, 0.6, 0.2, -0.4, 0.3])
    plan = mods["gto.utils"].interpolate_waypoints(np.stack([qc, qg]), 12, robot.ndof).T
    base = [0.05, -0.02, 0.1]
    cost, dist = robot.compute_plan_cost(plan, field, base)
    out["plan"], out["plan_base"], out["plan_cost"], out["plan_dist"] = plan, np.array(base), cost, dist
    pw, _ = robot.compute_fk_surface_points(qc)
    out["fk_points_qc"] = pw
    # setup_points_field on a synthetic cloud
    cloud = rng.uniform([0.2, -0.5, 0.0], [0.93, 0.61, 0.47], size=(500, 3))
    robot.setup_points_field(cloud)
    out["pf_cloud_min"], out["pf_cloud_max"] = cloud.min(0), cloud.max(0)
    out["pf_origin"], out["pf_shape"] = robot.origin.copy(), np.array(robot.field_shape)

    # DepthPointCloud: camera looking down at a table with a box on it
    DPC = mods["mesh_to_sdf.depth_point_cloud"].DepthPointCloud
    Hh, Ww = 48, 64
    Kc = np.array([[60.0, 0, Ww / 2], [0, 60.0, Hh / 2], [0, 0, 1]])
    depth = np.full((Hh, Ww), 1.0, dtype=np.float32)
    depth[16:32, 24:40] = 0.8
    cam = np.eye(4)
    cam[:3, :3] = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]])  # looking along -z
    cam[:3, 3] = [0.5, 0.0, 1.0]
    dpc = DPC(depth, Kc, cam, target_mask=None, threshold=1.5)
    qp = rng.uniform([0.0, -0.5, -0.1], [1.0, 0.5, 0.4], size=(300, 3))
    out["dpc_depth"], out["dpc_K"], out["dpc_cam"], out["dpc_query"] = depth, Kc, cam, qp
    out["dpc_points"] = dpc.points
    out["dpc_sdf"] = dpc.get_sdf(qp)
    out["dpc_cost"] = dpc.get_sdf_cost(qp, epsilon=0.02)
    np.savez_compressed(os.path.join(OUT, "ref_field.npz"), **out)
    print("ref_field.npz ws", out["ws_shape"], "pf", out["pf_shape"], "plan cost", cost, dist)


def gen_seed(mods, rng):
    iw = mods["gto.utils"].interpolate_waypoints
    qc = np.array([0.0, -1.285, 0, -2.356, 0.0, 1.571, 0.785, 0.0, 0.0])
    qg = qc + rng.normal(0, 0.5, size=9)
    out = {"qc": qc, "qg": qg}
    for T in (30, 50):
        out[f"cubic_T{T}"] = iw(np.stack([qc, qg]), T, 9)
        out[f"linear_T{T}"] = iw(np.stack([qc, qg]), T, 9, mode="linear")
    mid = 0.5 * (qc + qg) + 0.1
    out["mid"] = mid
    out["cubic3_T30"] = iw(np.stack([qc, mid, qg]), 30, 9)
    np.savez_compressed(os.path.join(OUT, "ref_seed.npz"), **out)
    print("ref_seed.npz")


def gen_stored_plans(ref):
    files = {
        "panda_tabletop": "GTO_scenereplica_panda_tabletop_24-02-06_T180750.json",
        "panda_shelf": "GTO_scenereplica_panda_shelf_24-02-06_T192709.json",
        "fetch_tabletop": "GTO_scenereplica_fetch_tabletop_24-02-06_T181818.json",
        "fetch_shelf": "GTO_scenereplica_fetch_shelf_24-02-06_T205216.json",
    }
    out = {}
    for key, fn in files.items():
        d = json.load(open(os.path.join(ref, "examples", "results_iros2024", fn)))
        plans, times = [], []
        for scene in sorted(d.keys(), key=lambda s: int(s)):
            for order in sorted(d[scene].keys()):
                for obj in sorted(d[scene][order].keys()):
                    e = d[scene][order][obj]
                    if e.get("plan") is not None and len(plans) < 16:
                        plans.append(np.array(e["plan"], dtype=np.float64))
                        times.append(float(e["planning_time"]))
        out[key] = np.stack(plans)
        out[key + "_planning_time"] = np.array(times)
    np.savez_compressed(os.path.join(OUT, "ref_stored_plans.npz"), **out)
    print("ref_stored_plans.npz", {k: v.shape for k, v in out.items()})


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    os.makedirs(OUT, exist_ok=True)
    mods = load_reference(ref)
    rng = np.random.default_rng(20240206)
    gen_fk(ref, mods, rng)
    gen_sdf(mods, rng)
    gen_field(ref, mods, rng)
    gen_seed(mods, rng)
    gen_stored_plans(ref)


if __name__ == "__main__":
    main()
