"""SURVEY.md section 8(f) row 2 on the GPU: DepthPointCloud.get_sdf / get_sdf_cost through libgto_b200 (k_cloud_query) against
(1) the outputs of the reference's own DepthPointCloud stored in tests/golden/ref_field.npz and (2) the reference algorithm
(oracle/dpc_oracle.py: scikit-learn KD-tree) on a larger synthetic depth image.  Run with -m gpu."""
import os
import time

import numpy as np
import pytest

from mesh_to_sdf.depth_point_cloud import DepthPointCloud
from dpc_oracle import KDTreeDepthPointCloud

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_cloud_query_matches_reference_run():
    z = np.load(os.path.join(GOLDEN, "ref_field.npz"))
    dpc = DepthPointCloud(z["dpc_depth"], z["dpc_K"], z["dpc_cam"], target_mask=None, threshold=1.5)
    np.testing.assert_allclose(dpc.points, z["dpc_points"], atol=1e-12)  # back-projection on the device (k_cloud_backproject)
    sdf = dpc.get_sdf(z["dpc_query"])
    assert sdf.dtype == np.float32 and sdf.shape == z["dpc_sdf"].shape
    np.testing.assert_array_equal(np.sign(sdf), np.sign(z["dpc_sdf"]))
    np.testing.assert_allclose(sdf, z["dpc_sdf"], rtol=0, atol=2e-6)  # float32 search vs float64 KD-tree cast to float32
    np.testing.assert_allclose(dpc.get_sdf_cost(z["dpc_query"], epsilon=0.02), z["dpc_cost"], rtol=0, atol=2e-6)
    assert (z["dpc_sdf"] < 0).any() and (z["dpc_cost"] > 0).any()


def _scene_depth(H=120, W=160):
    """Camera 1 m above a table looking straight down; a box and a slanted plane segment on the table."""
    f = 140.0
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.0]])
    cam = np.eye(4)
    cam[:3, :3] = np.array([[1.0, 0, 0], [0, -1, 0], [0, 0, -1]])  # optical axis = -z world
    cam[:3, 3] = [0.5, 0.0, 1.0]
    v, u = np.mgrid[0:H, 0:W]
    depth = np.full((H, W), 1.0, np.float32)
    depth[40:80, 50:100] = 0.85  # box top
    depth[10:30, 20:140] = (0.95 - 0.001 * (u[10:30, 20:140] - 20)).astype(np.float32)  # ramp
    depth[100:, :10] = 0.0  # invalid pixels
    return depth, K, cam


def test_cloud_query_matches_kdtree_backend_on_a_grid():
    depth, K, cam = _scene_depth()
    mask = np.zeros(depth.shape, np.uint8)
    mask[45:60, 55:70] = 1  # target object pixels are left out of the cloud
    for tm in (None, mask):
        gpu = DepthPointCloud(depth, K, cam, target_mask=tm, threshold=1.5)
        cpu = KDTreeDepthPointCloud(depth, K, cam, target_mask=tm, threshold=1.5)
        assert gpu.points.shape == cpu.points.shape
        np.testing.assert_allclose(gpu.points, cpu.points, rtol=0, atol=1e-14)
    n = 56
    g = np.stack(np.meshgrid(np.linspace(0.0, 1.0, n), np.linspace(-0.5, 0.5, n), np.linspace(-0.2, 0.8, n), indexing="ij"), axis=-1).reshape(-1, 3)
    t0 = time.time(); s_gpu = gpu.get_sdf(g); t1 = time.time(); s_cpu = cpu.get_sdf(g); t2 = time.time()
    print(f"cloud {gpu.points.shape[0]} points x {g.shape[0]} queries: GPU kernel {gpu.last_kernel_ms:.2f} ms (call {1e3*(t1-t0):.1f} ms), KD-tree {1e3*(t2-t1):.1f} ms")
    flip = np.sign(s_gpu) != np.sign(s_cpu)
    assert flip.mean() < 1e-4  # a query projecting exactly onto a pixel border may fall on either side
    np.testing.assert_allclose(np.abs(s_gpu), np.abs(s_cpu), rtol=0, atol=2e-6)
    c_gpu, c_cpu = gpu.get_sdf_cost(g, epsilon=0.02), cpu.get_sdf_cost(g, epsilon=0.02)
    ok = ~flip
    np.testing.assert_allclose(c_gpu[ok], c_cpu[ok], rtol=0, atol=2e-6)
    assert (c_cpu > 0).mean() > 0.05 and (s_cpu < 0).any()
    # is_outside on the device (mode 2 of gto_cloud_query) against the restated reference test
    vis_g, vis_c = gpu.is_outside(g), cpu.is_outside(g)
    assert vis_g.dtype == bool and (vis_g != vis_c).mean() < 1e-4 and (~vis_c).any()
    # the pruned search visits every tile that could hold a nearer point: bit-identical to the brute-force kernel
    os.environ["GTO_CLOUD_BRUTE"] = "1"
    try:
        s_brute, c_brute = gpu.get_sdf(g), gpu.get_sdf_cost(g, epsilon=0.02)
        print(f"brute-force kernel {gpu.last_kernel_ms:.2f} ms")
    finally:
        os.environ.pop("GTO_CLOUD_BRUTE")
    np.testing.assert_array_equal(s_gpu, s_brute)
    np.testing.assert_array_equal(c_gpu, c_brute)
    # ragged sizes: fewer queries than one block, a non-multiple of the tile
    for m in (1, 7, 1500):
        np.testing.assert_allclose(np.abs(gpu.get_sdf(g[:m])), np.abs(s_cpu[:m]), rtol=0, atol=2e-6)


def test_plan_collision_audit_matches_kdtree_backend(tmp_path):
    """Plan post-check (examples/pybullet_evaluate_plans.py:219-237): all knots of a plan in one GPU query, same per-knot counts as
    the reference's per-knot KD-tree loop."""
    from gto.gto_models import GTORobotModel
    from gto.utils import plan_collision_audit
    from test_compat_api import _write_box

    _write_box(tmp_path)
    robot = GTORobotModel(str(tmp_path), urdf_filename=str(tmp_path / "arm3.urdf"), time_derivs=[0, 1], param_joints=["slide"],
                          collision_link_names=["base", "l1", "l2", "tool"], sample_point_count=64, seed=2)
    depth, K, cam = _scene_depth()
    gpu = DepthPointCloud(depth, K, cam, threshold=1.5)
    cpu = KDTreeDepthPointCloud(depth, K, cam, threshold=1.5)
    T = 12
    plan = np.stack([np.linspace(-2.0, 2.0, T), np.linspace(0.0, 6.0, T), np.full(T, 0.01)])  # the arm starts inside the box on the table
    base = np.array([0.2, 0.0, 0.12])
    hit_g, first_g, cnt_g = plan_collision_audit(robot, plan, gpu, base)
    # the reference's loop, knot by knot
    cnt_c = np.array([(cpu.get_sdf(robot.compute_fk_surface_points(plan[:, i])[0] + base) < 0).sum() for i in range(T)])
    assert np.abs(cnt_g - cnt_c).max() <= 1  # a point projecting exactly onto a pixel border may fall on either side
    assert 0 < cnt_c[0] <= 5 < cnt_c[1] and cnt_c[2:].max() == 0  # knot 0 touches (4 points), knot 1 collides (6), the rest is free
    assert hit_g and first_g == int(np.flatnonzero(cnt_c > 5)[0]) == 1
    hit0, first0, _ = plan_collision_audit(robot, plan[:, :1] * 0, gpu, np.array([0.3, 0.0, 0.6]))
    assert (cpu.get_sdf(robot.compute_fk_surface_points(plan[:, 0] * 0)[0] + np.array([0.3, 0.0, 0.6])) < 0).sum() <= 5 and not hit0 and first0 == -1
