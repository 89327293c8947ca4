"""ndarray-backed stand-in for ``casadi.DM`` -- only the numeric behaviour the grasp-trajectory
callers use (``gto/gto_planner.py:27-28,152-158,194,215-219``; ``examples/pybullet_api.py``):
construction from lists/arrays (1-D input becomes a column, like CasADi), ``DM.ones/zeros/eye``,
``@``, arithmetic, slicing (read and write), ``.toarray()``/``.full()``, ``.shape``, ``.T``.

The symbolic ``MX``/``SX`` API of CasADi is deliberately NOT reproduced: the NLP is no longer
built symbolically."""
from __future__ import annotations

import numpy as np


def _as2d(x) -> np.ndarray:
    if isinstance(x, DM):
        return x._a
    a = np.array(x, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    elif a.ndim > 2:
        raise ValueError("DM is at most two-dimensional")
    return a


class DM:
    __array_priority__ = 100.0

    def __init__(self, *args):
        if len(args) == 0:
            self._a = np.zeros((0, 0))
        elif len(args) == 1:
            self._a = _as2d(args[0]).copy()
        elif len(args) == 2:
            self._a = np.zeros((int(args[0]), int(args[1])))
        else:
            raise TypeError("DM(x) or DM(rows, cols)")

    # -- constructors ------------------------------------------------------------------
    @staticmethod
    def zeros(m=1, n=1):
        return DM(np.zeros((int(m), int(n))))

    @staticmethod
    def ones(m=1, n=1):
        return DM(np.ones((int(m), int(n))))

    @staticmethod
    def eye(n):
        return DM(np.eye(int(n)))

    # -- conversions -------------------------------------------------------------------
    def toarray(self, simplify: bool = False):
        a = self._a.copy()
        if simplify:
            if a.size == 1:
                return float(a.reshape(-1)[0])
            if 1 in a.shape:
                return a.reshape(-1)
        return a

    def full(self):
        return self._a.copy()

    def __array__(self, dtype=None, copy=None):
        return self._a.astype(dtype) if dtype is not None else self._a

    def __float__(self):
        if self._a.size != 1:
            raise TypeError("only 1x1 DM converts to float")
        return float(self._a.reshape(-1)[0])

    def __int__(self):
        return int(float(self))

    # -- shape -------------------------------------------------------------------------
    @property
    def shape(self):
        return self._a.shape

    def size1(self):
        return self._a.shape[0]

    def size2(self):
        return self._a.shape[1]

    def numel(self):
        return self._a.size

    @property
    def T(self):
        return DM(self._a.T)

    def reshape(self, shape):
        # CasADi reshapes column-major
        return DM(self._a.reshape(shape, order="F"))

    def __len__(self):
        return self._a.shape[0]

    # -- indexing ----------------------------------------------------------------------
    @staticmethod
    def _key(k):
        if isinstance(k, DM):
            return k._a.reshape(-1).astype(np.int64)
        return k

    def _norm(self, key):
        if isinstance(key, tuple):
            r, c = (self._key(k) for k in key)
            if isinstance(r, (list, np.ndarray)) and isinstance(c, (list, np.ndarray)):
                return np.ix_(np.asarray(r).reshape(-1), np.asarray(c).reshape(-1))
            return (r, c)
        key = self._key(key)
        return key

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            # linear (column-major) indexing like CasADi
            flat = self._a.reshape(-1, order="F")[self._norm(key)]
            return DM(np.asarray(flat, dtype=np.float64).reshape(-1, 1))
        sub = self._a[self._norm(key)]
        sub = np.asarray(sub, dtype=np.float64)
        if sub.ndim == 0:
            sub = sub.reshape(1, 1)
        elif sub.ndim == 1:
            r, c = key
            # a scalar row index with a slice of columns gives a row, otherwise a column
            sub = sub.reshape(1, -1) if np.isscalar(r) or isinstance(r, (int, np.integer)) else sub.reshape(-1, 1)
        return DM(sub)

    def __setitem__(self, key, value):
        v = value._a if isinstance(value, DM) else np.asarray(value, dtype=np.float64)
        if not isinstance(key, tuple):
            flat = self._a.reshape(-1, order="F")
            flat[self._norm(key)] = np.asarray(v).reshape(-1) if np.size(v) > 1 else float(np.asarray(v).reshape(-1)[0])
            self._a = flat.reshape(self._a.shape, order="F")
            return
        idx = self._norm(key)
        target = self._a[idx]
        if np.ndim(target) == 0:
            self._a[idx] = float(np.asarray(v).reshape(-1)[0])
        else:
            self._a[idx] = np.asarray(v).reshape(np.shape(target)) if np.size(v) == np.size(target) else v

    # -- arithmetic --------------------------------------------------------------------
    @staticmethod
    def _val(o):
        if isinstance(o, DM):
            return o._a
        a = np.asarray(o, dtype=np.float64)
        return a.reshape(-1, 1) if a.ndim == 1 else a

    def __matmul__(self, o):
        return DM(self._a @ self._val(o))

    def __rmatmul__(self, o):
        return DM(self._val(o) @ self._a)

    def __add__(self, o):
        return DM(self._a + self._val(o))

    __radd__ = __add__

    def __sub__(self, o):
        return DM(self._a - self._val(o))

    def __rsub__(self, o):
        return DM(self._val(o) - self._a)

    def __mul__(self, o):
        return DM(self._a * self._val(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return DM(self._a / self._val(o))

    def __rtruediv__(self, o):
        return DM(self._val(o) / self._a)

    def __neg__(self):
        return DM(-self._a)

    def __pow__(self, e):
        return DM(self._a**e)

    def __lt__(self, o):
        return DM((self._a < self._val(o)).astype(np.float64))

    def __le__(self, o):
        return DM((self._a <= self._val(o)).astype(np.float64))

    def __gt__(self, o):
        return DM((self._a > self._val(o)).astype(np.float64))

    def __ge__(self, o):
        return DM((self._a >= self._val(o)).astype(np.float64))

    def __bool__(self):
        if self._a.size != 1:
            raise ValueError("truth value of a non-scalar DM is ambiguous")
        return bool(self._a.reshape(-1)[0])

    def __repr__(self):
        return f"DM({np.array2string(self._a, precision=6)})"

    __str__ = __repr__


# -- free functions mirroring the casadi namespace pulled in by ``from casadi import *`` ---------
def vertcat(*xs):
    xs = [_as2d(x) for x in xs if np.size(_as2d(x)) > 0]
    return DM(np.vstack(xs)) if xs else DM()


def horzcat(*xs):
    xs = [_as2d(x) for x in xs if np.size(_as2d(x)) > 0]
    return DM(np.hstack(xs)) if xs else DM()


def vec(x):
    return DM(_as2d(x).reshape(-1, 1, order="F"))


def diag(x):
    a = _as2d(x)
    if 1 in a.shape:
        return DM(np.diag(a.reshape(-1)))
    return DM(np.diag(a).reshape(-1, 1))


def linspace(a, b, n):
    return DM(np.linspace(float(a), float(b), int(n)).reshape(-1, 1))


def sumsqr(x):
    a = _as2d(x)
    return DM(np.sum(a * a))


def sum1(x):
    return DM(np.sum(_as2d(x), axis=0, keepdims=True))


def sum2(x):
    return DM(np.sum(_as2d(x), axis=1, keepdims=True))


def mmin(x):
    return DM(np.min(_as2d(x)))


def mmax(x):
    return DM(np.max(_as2d(x)))


def norm_2(x):
    return DM(np.linalg.norm(_as2d(x).reshape(-1)))


norm_fro = norm_2


def _ufunc(f):
    def g(x, *rest):
        return DM(f(_as2d(x), *[_as2d(r) for r in rest]))

    return g


sin = _ufunc(np.sin)
cos = _ufunc(np.cos)
tan = _ufunc(np.tan)
sqrt = _ufunc(np.sqrt)
floor = _ufunc(np.floor)
ceil = _ufunc(np.ceil)
fabs = _ufunc(np.abs)
fmin = _ufunc(np.minimum)
fmax = _ufunc(np.maximum)
atan2 = _ufunc(np.arctan2)
acos = _ufunc(np.arccos)
asin = _ufunc(np.arcsin)
exp = _ufunc(np.exp)
log = _ufunc(np.log)
