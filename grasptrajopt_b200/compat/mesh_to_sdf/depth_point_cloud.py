"""``DepthPointCloud`` on the GPU: depth image -> world point cloud -> signed distance (nearest cloud point, negative where the
query is hidden behind the visible surface) -> CHOMP-style cost field.  Producer of the voxel fields the hot path consumes
(SURVEY.md section 8(f) row 2).

Interface of the reference class (``mesh_to_sdf/depth_point_cloud.py:9-142``: constructor arguments, the attributes ``points``,
``depth``, ``intrinsic_matrix``, ``camera_pose``, ``target_mask``, ``width``, ``height``, ``threshold`` and the methods below) with
every computation in ``libgto_b200.so``:

    back-projection            ``gto_cloud_backproject``  (k_cloud_backproject, one thread per pixel, float64)
    get_sdf / get_sdf_cost     ``gto_cloud_query`` mode 0 / 1  (k_cloud_query_pruned: exact nearest neighbour over Morton-sorted tiles,
                               visibility sign by float64 projection into the depth image, cost transform fused)
    is_outside                 ``gto_cloud_query`` mode 2  (the visibility test alone)

There is no CPU path: without a CUDA device the constructor raises (``GtoError``).  The KD-tree restatement of the reference's
algorithm that the tests check this class against lives in ``oracle/dpc_oracle.py`` (test infrastructure).  The ``vis=True``
branch of the reference (a pyrender viewer) is out of scope.
"""
from __future__ import annotations

import numpy as np


class DepthPointCloud:
    def __init__(self, depth, intrinsic_matrix, camera_pose, target_mask=None, threshold=1.5, device=0):
        from gto.b200_solver import get_context

        self.depth = depth
        self.intrinsic_matrix = intrinsic_matrix
        self.camera_pose = camera_pose
        self.target_mask = target_mask
        self.height, self.width = depth.shape[0], depth.shape[1]
        self.threshold = threshold
        self.device = device
        self.last_kernel_ms = None
        self._ctx = get_context(device)
        self._cam_inv = np.linalg.inv(np.asarray(camera_pose, dtype=np.float64))
        self._depth32 = np.ascontiguousarray(depth, dtype=np.float32)
        self.points = self._ctx.cloud_backproject(self._depth32, intrinsic_matrix, camera_pose, threshold, target_mask)
        self._make_resident()

    def _make_resident(self):
        # one cloud is resident per solver context; the last DepthPointCloud that used the context owns it
        self._ctx.cloud_set(self.points)
        self._ctx._cloud_owner = self

    def _query(self, query_points, mode, epsilon=0.02, w_inside=1):
        if mode != 2 and getattr(self._ctx, "_cloud_owner", None) is not self:
            self._make_resident()
        out, self.last_kernel_ms = self._ctx.cloud_query(np.asarray(query_points, dtype=np.float64), self._depth32, self.intrinsic_matrix, self._cam_inv,
                                                         mode, epsilon, w_inside)
        return out

    def get_random_surface_points(self, count):
        return self.points[np.random.choice(len(self.points), count)]

    def is_outside(self, points):
        """True where a point is visible from the camera or projects outside the image."""
        return self._query(points, 2) > 0.5

    def get_sdf(self, query_points):
        return self._query(query_points, 0)

    def get_sdf_cost(self, query_points, epsilon=0.02, w_inside=1, vis=False):
        if vis:
            raise NotImplementedError("the pyrender viewer of the reference (vis=True) is not part of the B200 build")
        return self._query(query_points, 1, epsilon, w_inside)

    def get_sdf_in_batches(self, query_points, batch_size=1000000):
        query_points = np.asarray(query_points)
        return np.concatenate([self.get_sdf(query_points[i : i + batch_size]) for i in range(0, len(query_points), batch_size)])
