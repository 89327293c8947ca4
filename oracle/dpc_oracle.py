"""CPU restatement of the reference's ``DepthPointCloud`` (``mesh_to_sdf/depth_point_cloud.py:9-142``) with a scikit-learn KD-tree.
TEST INFRASTRUCTURE ONLY: the checker for ``grasptrajopt_b200/compat/mesh_to_sdf/depth_point_cloud.py`` (tests/test_gpu_cloud.py,
tools/bench_rows_f.py); nothing in the product package imports it.  Pinned against the outputs of the reference's own class
run here (``tests/golden/ref_field.npz``, generator ``oracle/make_golden.py``) in tests/test_oracle_golden.py.

  cloud          pixels with 0 < depth < threshold outside the target mask, X = depth * K^-1 [u v 1]^T, world = R X + t   (:15-19,32-52)
  get_sdf        nearest-neighbour distance (KD-tree), negated where the query is hidden behind the visible surface     (:57-62)
  is_outside     project with K into the image through the inverse pose; outside the viewport or nearer than the depth   (:127-142)
  get_sdf_cost   d < 0: w_inside (-d + eps/2);  0 < d < eps: (d - eps)^2 / (2 eps);  else 0                             (:65-91)
"""
import numpy as np
from sklearn.neighbors import KDTree


def backproject(depth, K, camera_pose, threshold=1.5, target_mask=None):
    d = np.asarray(depth, dtype=np.float32)
    H, W = d.shape
    keep = (d > 0) & (d < threshold)
    if target_mask is not None:
        keep &= np.asarray(target_mask).reshape(H, W) == 0
    v, u = np.nonzero(keep)  # row-major pixel order
    rays = np.linalg.inv(np.asarray(K, dtype=np.float64)) @ np.stack([u, v, np.ones_like(u)]).astype(np.float64)
    cam = rays * d[v, u].astype(np.float64)
    pose = np.asarray(camera_pose, dtype=np.float64)
    return (pose[:3, :3] @ cam + pose[:3, 3:4]).T


class KDTreeDepthPointCloud:
    def __init__(self, depth, intrinsic_matrix, camera_pose, target_mask=None, threshold=1.5):
        self.depth, self.K, self.pose = np.asarray(depth), np.asarray(intrinsic_matrix, dtype=np.float64), np.asarray(camera_pose, dtype=np.float64)
        self.height, self.width = self.depth.shape
        self.points = backproject(depth, intrinsic_matrix, camera_pose, threshold, target_mask)
        self.tree = KDTree(self.points)

    def is_outside(self, points):
        inv = np.linalg.inv(self.pose)
        cam = np.asarray(points, dtype=np.float64) @ inv[:3, :3].T + inv[:3, 3]
        with np.errstate(divide="ignore", invalid="ignore"):
            proj = cam @ self.K.T
            px = (proj[:, :2] / proj[:, 2:3])
            ok = np.isfinite(px).all(axis=1) & (np.abs(px) < 2.0e9).all(axis=1)
            pix = np.where(ok[:, None], px, -1.0).astype(np.int64)  # astype(int): truncation toward zero
        inside_image = ok & (pix[:, 0] >= 0) & (pix[:, 1] >= 0) & (pix[:, 0] < self.width) & (pix[:, 1] < self.height)
        out = np.ones(len(cam), dtype=bool)
        out[inside_image] = cam[inside_image, 2] < self.depth[pix[inside_image, 1], pix[inside_image, 0]]
        return out

    def get_sdf(self, query_points):
        dist = self.tree.query(np.asarray(query_points, dtype=np.float64))[0].astype(np.float32).reshape(-1)
        dist[~self.is_outside(query_points)] *= -1
        return dist

    def get_sdf_cost(self, query_points, epsilon=0.02, w_inside=1):
        d = self.get_sdf(query_points)
        cost = np.zeros_like(d)
        neg = d < 0
        cost[neg] = w_inside * (-d[neg] + epsilon / 2)
        band = (d > 0) & (d < epsilon)
        cost[band] = np.square(d[band] - epsilon) / (2 * epsilon)
        return cost
