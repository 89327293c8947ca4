"""Numeric stand-in for the subset of ``casadi`` that the reference's *numeric* code paths touch.

TEST INFRASTRUCTURE (used only by ``oracle/make_golden.py``).  CasADi is not installable offline, so
the reference's own Python (``optas/spatialmath.py``, ``optas/models.py`` FK, ``gto/sdf_callback.py``,
``gto/gto_models.py`` numeric helpers) is executed on top of this ndarray-backed ``DM``.  Only matrix
plumbing lives here (concatenate, slice, ``@``, sin/cos, norm); every convention being pinned
(rotation order, Rodrigues form, joint indexing, voxel indexing, finite-difference stencils) is executed
from the reference's files where they lie.  Nothing symbolic is emulated: ``MX``/``SX`` exist as empty
types for ``isinstance`` checks only.
"""
import numpy as np  # re-exported: the reference uses ``cs.np``


def _a(x):
    if isinstance(x, DM):
        return x.a
    v = np.array(x, dtype=np.float64)
    if v.ndim == 0:
        return v.reshape(1, 1)
    if v.ndim == 1:
        return v.reshape(-1, 1)
    return v


class DM:
    __array_priority__ = 1000

    def __init__(self, *args):
        if len(args) == 0:
            self.a = np.zeros((0, 1))
        elif len(args) == 1:
            self.a = _a(args[0]).copy()
        else:
            self.a = np.zeros((int(args[0]), int(args[1])))

    @staticmethod
    def eye(n):
        return DM(np.eye(n))

    @staticmethod
    def zeros(*s):
        s = s if len(s) > 1 else (s[0], 1) if len(s) == 1 else (1, 1)
        return DM(np.zeros(s))

    @staticmethod
    def ones(*s):
        s = s if len(s) > 1 else (s[0], 1) if len(s) == 1 else (1, 1)
        return DM(np.ones(s))

    @property
    def shape(self):
        return self.a.shape

    @property
    def T(self):
        return DM(self.a.T)

    def toarray(self):
        return self.a.copy()

    def full(self):
        return self.a.copy()

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __float__(self):
        return float(self.a.reshape(-1)[0])

    def __getitem__(self, k):
        if not isinstance(k, tuple):
            return DM(self.a.reshape(-1, order="F")[k])
        r = self.a[k]
        if np.ndim(r) == 1:
            r = r.reshape(1, -1) if isinstance(k[0], (int, np.integer)) else r.reshape(-1, 1)
        return DM(r)

    def __setitem__(self, k, v):
        self.a[k] = np.asarray(_a(v)).reshape(np.shape(self.a[k]))

    def _b(self, o, f):
        return DM(f(self.a, _a(o)))

    def __matmul__(self, o):
        return DM(self.a @ _a(o))

    def __rmatmul__(self, o):
        return DM(_a(o) @ self.a)

    def __add__(self, o):
        return self._b(o, np.add)

    __radd__ = __add__

    def __sub__(self, o):
        return self._b(o, np.subtract)

    def __rsub__(self, o):
        return DM(_a(o) - self.a)

    def __mul__(self, o):
        return self._b(o, np.multiply)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._b(o, np.divide)

    def __neg__(self):
        return DM(-self.a)


class SX:  # isinstance targets only
    pass


class MX:
    pass


def vertcat(*xs):
    xs = [_a(x) for x in xs]
    xs = [x for x in xs if x.size]
    return DM(np.vstack(xs))


def horzcat(*xs):
    xs = [_a(x) for x in xs]
    xs = [x for x in xs if x.size]
    return DM(np.hstack(xs))


def vertsplit(x):
    return [DM(r.reshape(1, -1)) for r in _a(x)]


def horzsplit(x):
    return [DM(c.reshape(-1, 1)) for c in _a(x).T]


def vec(x):
    return DM(_a(x).reshape(-1, 1, order="F"))


def sin(x):
    return DM(np.sin(_a(x)))


def cos(x):
    return DM(np.cos(_a(x)))


def norm_fro(x):
    return DM(np.linalg.norm(_a(x)))


def floor(x):
    return DM(np.floor(_a(x)))


def fmax(x, y):
    return DM(np.maximum(_a(x), _a(y)))


def fmin(x, y):
    return DM(np.minimum(_a(x), _a(y)))


class Sparsity:
    @staticmethod
    def dense(m, n=1):
        return (m, n)


class Callback:
    """``casadi.Callback`` protocol as used by ``gto/sdf_callback.py``: ``construct`` runs ``init``;
    calling the object evaluates ``eval`` column by column on a 3-by-n input."""

    def __init__(self):
        pass

    def construct(self, name, opts={}):
        self.init() if hasattr(self, "init") else None

    def __call__(self, *args):
        x = _a(args[0])
        outs = [self.eval([DM(x[:, i : i + 1])] + [None] * (self.get_n_in() - 1)) for i in range(x.shape[1])]
        return outs


class Function:  # annotation target only; nothing symbolic is emulated
    def __init__(self, *a, **k):
        raise NotImplementedError("symbolic casadi.Function is not emulated")
