"""Build the robot tables (joint table + seeded surface point sets) the benchmark and the GPU tests use.

Inputs are the reference checkout's URDFs, YAML configs and visual meshes (``data/robots``, ``data/configs``);
they are read where they lie and only the derived tables are written (a few tens of KB each), so the
GPU box -- which has no ``/root/reference`` -- can run the BASELINE configs:

  panda_c2.npz    Panda, 7 optimised joints, 12 collision links, P=2000   (configs C1, C2, C5)
  fetch8_c3.npz   Fetch with torso_lift_joint optimised (8), 10 links, P=4000  (C3)
  fetch10_c4.npz  Fetch (7 arm joints, shipped YAML) + planar base (x, y, yaw) virtual joints (10), P=4000 (C4)
  panda_small.npz Panda, P=12x32 (fast parity tests)

Run:  python grasptrajopt_b200/assets/make_assets.py [/root/reference]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import grasptrajopt_b200 as pkg  # noqa: E402

pkg.install_compat()
from gto.gto_models import GTORobotModel  # noqa: E402
from gto.utils import load_yaml  # noqa: E402
from grasptrajopt_b200.robot_table import prepend_planar_base  # noqa: E402


def even_counts(total, nlinks):
    base = total // nlinks
    counts = [base] * nlinks
    for i in range(total - base * nlinks):
        counts[i] += 1
    return counts


def build(ref, robot, total_points, drop_param=(), seed=0):
    cfg = load_yaml(os.path.join(ref, "data", "configs", f"{robot}.yaml"))["robot_cfg"]
    model_dir = os.path.join(ref, "data", "robots", cfg["robot_name"])
    params = [j for j in cfg["param_joints"] if j not in drop_param]
    links = cfg["collision_link_names"]
    per = max(even_counts(total_points, len(links)))
    model = GTORobotModel(model_dir, urdf_filename=os.path.join(ref, cfg["urdf_robot_path"]), time_derivs=[0, 1],
                          param_joints=params, collision_link_names=links, sample_point_count=per, seed=seed)
    counts = even_counts(total_points, len(model.surface_pc_map))
    for (name, pc), c in zip(model.surface_pc_map.items(), counts):
        pc.points, pc.normals = pc.points[:c], pc.normals[:c]
    return model.to_table(cfg["link_ee"], cfg["link_gripper"]), cfg


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    t, cfg = build(ref, "panda", 2000)
    t.save(os.path.join(HERE, "panda_c2.npz"))
    print("panda_c2", t.nopt, t.nmov, t.nlinks, t.npoints, t.grip_pt_count)
    t, _ = build(ref, "panda", 12 * 32)
    t.save(os.path.join(HERE, "panda_small.npz"))
    t8, _ = build(ref, "fetch", 4000, drop_param=("torso_lift_joint",))
    t8.save(os.path.join(HERE, "fetch8_c3.npz"))
    print("fetch8_c3", t8.nopt, t8.nmov, t8.nlinks, t8.npoints, t8.grip_pt_count)
    t7, _ = build(ref, "fetch", 4000)
    t10 = prepend_planar_base(t7)  # "10-DoF mobile" = 7 arm joints + planar base (x, y, yaw)
    t10.save(os.path.join(HERE, "fetch10_c4.npz"))
    print("fetch10_c4", t10.nopt, t10.nmov, t10.nlinks, t10.npoints)
    t7, _ = build(ref, "fetch", 10 * 32)
    t7.save(os.path.join(HERE, "fetch_small.npz"))


if __name__ == "__main__":
    main()
