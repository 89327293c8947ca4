"""Triangle-mesh reading (OBJ / STL) and seeded area-weighted surface sampling.

Replaces, for the hot path's inputs, what the reference gets from ``trimesh``:
``trimesh.load(filename)`` followed by ``mesh.sample(count, return_index=True)``
(``gto/gto_models.py:75-77``, ``mesh_to_sdf/surface_point_cloud.py:177-188``).
The reference sampler is *unseeded* (SURVEY.md Appendix C, Q2), so point sets can
never match it bit for bit; this sampler is seeded so that every parity test and
every C-ABI call sees explicit, reproducible point sets.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np


@dataclass
class TriMesh:
    vertices: np.ndarray  # [V,3] float64
    faces: np.ndarray  # [F,3] int64

    @property
    def triangles(self) -> np.ndarray:
        return self.vertices[self.faces]

    @property
    def face_normals(self) -> np.ndarray:
        tri = self.triangles
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        ln = np.linalg.norm(n, axis=1, keepdims=True)
        return n / np.where(ln > 0, ln, 1.0)

    @property
    def area_faces(self) -> np.ndarray:
        tri = self.triangles
        return 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)


def _load_obj(path: str) -> TriMesh:
    verts = []
    faces = []
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):  # fan triangulation of polygons
                    faces.append((idx[0], idx[k], idx[k + 1]))
    return TriMesh(np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(faces, dtype=np.int64).reshape(-1, 3))


def _load_stl(path: str) -> TriMesh:
    with open(path, "rb") as fh:
        data = fh.read()
    ntri = struct.unpack_from("<I", data, 80)[0] if len(data) >= 84 else 0
    if len(data) == 84 + 50 * ntri and ntri > 0:  # binary STL
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=ntri, offset=84)
        tri = rec["v"].astype(np.float64)
    else:  # ASCII STL
        pts = []
        for line in data.decode("utf-8", errors="ignore").splitlines():
            s = line.strip()
            if s.startswith("vertex"):
                p = s.split()
                pts.append((float(p[1]), float(p[2]), float(p[3])))
        tri = np.asarray(pts, dtype=np.float64).reshape(-1, 3, 3)
    verts = tri.reshape(-1, 3)
    faces = np.arange(verts.shape[0], dtype=np.int64).reshape(-1, 3)
    return TriMesh(verts, faces)


def load_mesh(path: str, scale=None) -> TriMesh:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".obj":
        mesh = _load_obj(path)
    elif ext == ".stl":
        mesh = _load_stl(path)
    else:
        raise ValueError(f"unsupported mesh format '{ext}' ({path}); convert to OBJ or STL")
    if mesh.faces.shape[0] == 0:
        raise ValueError(f"mesh {path} has no triangles")
    if scale is not None:
        mesh.vertices = mesh.vertices * np.asarray(scale, dtype=np.float64).reshape(1, 3)
    return mesh


def sample_surface(mesh: TriMesh, count: int, seed) -> tuple:
    """Area-weighted uniform samples on the mesh surface.

    Returns ``(points [count,3], normals [count,3])`` as float64; ``seed`` is anything
    ``np.random.default_rng`` accepts.
    """
    rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
    area = mesh.area_faces
    cdf = np.cumsum(area)
    fidx = np.searchsorted(cdf, rng.random(count) * cdf[-1], side="right")
    fidx = np.minimum(fidx, len(area) - 1)
    tri = mesh.triangles[fidx]
    u = rng.random((count, 1))
    v = rng.random((count, 1))
    flip = (u + v) > 1.0
    u = np.where(flip, 1.0 - u, u)
    v = np.where(flip, 1.0 - v, v)
    pts = tri[:, 0] + u * (tri[:, 1] - tri[:, 0]) + v * (tri[:, 2] - tri[:, 0])
    return pts, mesh.face_normals[fidx]


@dataclass
class SurfacePointCloud:
    """Same attribute names as ``mesh_to_sdf.surface_point_cloud.SurfacePointCloud``
    (``mesh_to_sdf/surface_point_cloud.py:31-36``) as far as the hot path reads them."""

    points: np.ndarray
    normals: np.ndarray
