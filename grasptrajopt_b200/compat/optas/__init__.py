"""CasADi-free ``optas`` namespace for the grasp-trajectory path.

The reference's ``optas/__init__.py:1-8`` re-exports all of CasADi plus its own modelling layer.
Here only the numeric surface used by ``gto`` and the example scripts exists:
``DM`` and a few free functions, ``spatialmath``, ``RobotModel``/``TaskModel``.  The symbolic
``OptimizationBuilder`` + ``CasADiSolver`` pair is replaced, for this one problem family, by
``gto.b200_solver.B200Solver`` (CUDA, through ``libgto_b200.so``)."""
from .dm import (DM, vertcat, horzcat, vec, diag, linspace, sumsqr, sum1, sum2, mmin, mmax, norm_2, norm_fro,
                 sin, cos, tan, sqrt, floor, ceil, fabs, fmin, fmax, atan2, acos, asin, exp, log)
from . import spatialmath
from .spatialmath import *  # noqa: F401,F403
from .spatialmath import pi, eps
from .models import RobotModel, TaskModel, Model, JointTypeNotSupported
import numpy as np

inf = float("inf")


def deg2rad(x):
    return DM((pi / 180.0) * np.asarray(DM(x)))


def rad2deg(x):
    return DM((180.0 / pi) * np.asarray(DM(x)))


def clip(x, lo, hi):
    return fmax(fmin(x, hi), lo)


def __getattr__(name):
    if name == "Visualizer":  # lazy: needs vtk (reference imports it eagerly, optas/visualize.py:4-6)
        from .visualize import Visualizer
        return Visualizer
    if name in ("OptimizationBuilder", "CasADiSolver", "OSQPSolver", "CVXOPTSolver", "ScipyMinimizeSolver", "MX", "SX"):
        raise AttributeError(
            f"optas.{name} is not part of the B200 build: the symbolic CasADi modelling layer is replaced by "
            "gto.b200_solver.B200Solver for the grasp-trajectory problem family (see INTEGRATION.md)")
    raise AttributeError(name)
