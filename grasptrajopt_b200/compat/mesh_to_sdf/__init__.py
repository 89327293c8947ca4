"""``mesh_to_sdf`` surface used by the grasp-trajectory path (reference ``mesh_to_sdf/__init__.py:7-21``).

Only ``surface_point_method='sample'`` exists here -- the one ``gto`` selects
(``gto/gto_models.py:76``).  The scan-based method needs an OpenGL renderer and is out of scope
(SURVEY.md section 2).  Unlike the reference (unseeded ``trimesh.sample``) the sampler takes a seed."""
from __future__ import annotations

from grasptrajopt_b200.meshio import TriMesh, load_mesh, sample_surface, SurfacePointCloud


def get_surface_point_cloud(mesh, surface_point_method="sample", bounding_radius=None, scan_count=100, scan_resolution=400,
                            sample_point_count=10000000, calculate_normals=True, seed=0):
    if not isinstance(mesh, TriMesh):
        raise TypeError("The mesh parameter must be a grasptrajopt_b200.meshio.TriMesh (use load_mesh).")
    if surface_point_method != "sample":
        raise ValueError("only surface_point_method='sample' is available in the B200 build (scan needs pyrender/OpenGL)")
    points, normals = sample_surface(mesh, sample_point_count, seed)
    return SurfacePointCloud(points=points, normals=normals if calculate_normals else None)
