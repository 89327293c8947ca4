"""``optas.spatialmath`` surface on NumPy (reference: ``optas/spatialmath.py``).

Thin DM-returning wrappers over ``grasptrajopt_b200.spatial`` so that user code written against
the reference (``from optas.spatialmath import standoff, rt2tr, rpy2r``) keeps working without
CasADi."""
from __future__ import annotations

import numpy as np

from grasptrajopt_b200 import spatial as _sp
from .dm import DM, _as2d

ArrayType = object
CasADiArrayType = DM
pi = _sp.pi
eps = _sp.eps


def arrayify_args(fun):
    """The reference decorator converts array-likes to CasADi arrays; here inputs are taken as-is."""
    return fun


def _v(x):
    return _as2d(x).reshape(-1)


def I3():
    return DM.eye(3)


def I4():
    return DM.eye(4)


def rotx(theta):
    return DM(_sp.rotx(float(theta)))


def roty(theta):
    return DM(_sp.roty(float(theta)))


def rotz(theta):
    return DM(_sp.rotz(float(theta)))


def rpy2r(rpy, opt="zyx"):
    return DM(_sp.rpy2r(_v(rpy), opt))


def angvec2r(theta, v):
    return DM(_sp.angvec2r(float(theta), _v(v)))


def r2t(R):
    return DM(_sp.r2t(_as2d(R)))


def rt2tr(R, t):
    return DM(_sp.rt2tr(_as2d(R), _v(t)))


def t2r(T):
    return DM(_as2d(T)[:3, :3])


def transl(T):
    return DM(_as2d(T)[:3, 3].reshape(3, 1))


def invt(T):
    return DM(_sp.invt(_as2d(T)))


def skew(v):
    return DM(_sp.skew(_v(v)))


def unit(v):
    return DM(_sp.unit(_v(v)).reshape(-1, 1))


def standoff(offset, axis="x"):
    return DM(_sp.standoff(float(offset), axis))


def vex(S):
    S = _as2d(S)
    if S.shape == (2, 2):
        return DM(0.5 * (S[1, 0] - S[0, 1]))
    return DM(0.5 * np.array([S[2, 1] - S[1, 2], S[0, 2] - S[2, 0], S[1, 0] - S[0, 1]]).reshape(3, 1))


class Quaternion:
    """Scalar-last quaternion (x, y, z, w) like ``optas.spatialmath.Quaternion`` (:303-458); numeric."""

    def __init__(self, x, y, z, w):
        self._q = np.array([float(x), float(y), float(z), float(w)])

    @staticmethod
    def fromrpy(rpy):
        return Quaternion.fromr(_sp.rpy2r(_v(rpy)))

    @staticmethod
    def fromr(R):
        w, x, y, z = _sp.mat2quat_wxyz(_as2d(R))
        return Quaternion(x, y, z, w)

    def getquat(self):
        return DM(self._q.reshape(4, 1))

    def getrpy(self):
        x, y, z, w = self._q
        r = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
        p = np.arcsin(np.clip(2 * (w * y - z * x), -1, 1))
        yw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
        return DM(np.array([r, p, yw]).reshape(3, 1))

    def inv(self):
        x, y, z, w = self._q
        n = float(self._q @ self._q)
        return Quaternion(-x / n, -y / n, -z / n, w / n)

    def __mul__(self, o):
        x1, y1, z1, w1 = self._q
        x2, y2, z2, w2 = o._q
        return Quaternion(
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
        )
