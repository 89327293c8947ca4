"""grasptrajopt_b200 -- Blackwell-native batched grasp-trajectory solver.

Layout:
  csrc/        CUDA kernels (sm_100a) + the C-ABI (``libgto_b200.so``; header ``include/gto_b200.h``)
  capi.py      ctypes binding of the C-ABI
  workloads.py, scenes.py     synthetic BASELINE configurations (C1..C5), scenes and cost fields
  robot_table.py, urdf.py, meshio.py, spatial.py, kinematics.py   host-side model/input preparation
  distributed.py, goalset.py  sharding, NCCL all-gather of results, goal-set arg-min
  run_reference.py            launcher that runs the reference's own scripts unchanged against compat/
  compat/      host-side mirror of the reference interface: top-level ``optas``, ``gto``, ``mesh_to_sdf``
"""
import os
import sys

__version__ = "0.1.0"

COMPAT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")


def install_compat() -> str:
    """Put the reference-compatible ``optas`` / ``gto`` / ``mesh_to_sdf`` packages on ``sys.path`` so
    that ``from gto.gto_planner import GTOPlanner`` resolves to the B200 implementation."""
    if COMPAT_DIR not in sys.path:
        sys.path.insert(0, COMPAT_DIR)
    return COMPAT_DIR
