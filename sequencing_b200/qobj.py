"""Dense stand-in for the subset of ``qutip.Qobj`` algebra the reference touches.

QuTiP is absent from this environment, and the only thing the reference needs
from it *above* the solve boundary is small-operator algebra: ``qeye/destroy/basis/tensor/
ket2dm/displace`` constructors, ``dag/expm/unit/ptrace/full/dims/isket/data.nnz`` and the
``expect/fidelity/tracedist`` helpers (163 call sites across ``sequencing/modes.py``,
``system.py``, ``gates/*.py``, ``benchmarking.py``).  Operators on this path are at most a few
thousand rows, built once per sweep, so a dense complex128 ndarray is the right container: it is
also exactly the layout the engine consumes (row-major ``[N, N]`` c128).
"""
from __future__ import annotations

import numbers
from functools import reduce
from typing import Iterable, List, Sequence

import numpy as np
import scipy.linalg as sla

__all__ = [
    "Qobj", "qeye", "identity", "destroy", "create", "num", "basis", "fock", "fock_dm",
    "ket2dm", "tensor", "displace", "coherent", "expect", "fidelity", "tracedist",
    "sigmax", "sigmay", "sigmaz", "isket", "isoper",
]


class _DataView:
    """Tiny shim for ``Qobj.data``: the reference only reads ``.nnz`` (system.py:396,420,444)."""

    __slots__ = ("_a",)

    def __init__(self, a):
        self._a = a

    @property
    def nnz(self) -> int:
        return int(np.count_nonzero(self._a))

    def toarray(self):
        return self._a.copy()


class Qobj:
    """Dense quantum object: ket ``[N,1]``, bra ``[1,N]`` or operator ``[N,N]``."""

    __array_priority__ = 1000  # numpy scalars defer to our __rmul__

    def __init__(self, inpt=None, dims=None, copy=True):
        if isinstance(inpt, Qobj):
            arr = inpt._a
            dims = dims or inpt.dims
        else:
            arr = np.asarray(inpt, dtype=complex)
        if arr.ndim == 0:
            arr = arr.reshape(1, 1)
        elif arr.ndim == 1:
            arr = arr.reshape(-1, 1)
        elif arr.ndim != 2:
            raise ValueError("Qobj data must be at most two-dimensional.")
        self._a = np.array(arr, dtype=complex, copy=True) if copy else arr
        if dims is None:
            dims = [[self._a.shape[0]], [self._a.shape[1]]]
        self.dims = [list(dims[0]), list(dims[1])]

    # -- structure -----------------------------------------------------------------------
    @property
    def shape(self):
        return self._a.shape

    @property
    def data(self):
        return _DataView(self._a)

    @property
    def isket(self) -> bool:
        return self._a.shape[1] == 1

    @property
    def isbra(self) -> bool:
        return self._a.shape[0] == 1 and self._a.shape[1] > 1

    @property
    def isoper(self) -> bool:
        return self._a.shape[0] == self._a.shape[1] and not self.isket

    @property
    def type(self) -> str:
        return "ket" if self.isket else "bra" if self.isbra else "oper"

    @property
    def isherm(self) -> bool:
        a = self._a
        if a.shape[0] != a.shape[1]:
            return False
        at = a.conj().T
        return bool(np.array_equal(a, at) or np.allclose(a, at, atol=1e-12, rtol=0))      # (exact test first: most operators are exactly Hermitian)

    def full(self, order="C", squeeze=False):
        out = np.array(self._a, order=order)
        return out.squeeze() if squeeze else out

    def copy(self):
        return Qobj(self._a, dims=self.dims)

    # -- algebra -------------------------------------------------------------------------
    def _wrap(self, arr, dims=None):
        return Qobj(arr, dims=dims or self.dims, copy=False)

    def __add__(self, other):
        if isinstance(other, Qobj):
            if other._a.shape != self._a.shape:
                raise TypeError("Incompatible quantum object shapes for +.")
            return self._wrap(self._a + other._a)
        if isinstance(other, numbers.Number):
            if other == 0:
                return self.copy()
            if self.isoper:
                return self._wrap(self._a + other * np.eye(self._a.shape[0]))
            return self._wrap(self._a + other)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        return self + (-1) * other

    def __rsub__(self, other):
        return (-1) * self + other

    def __neg__(self):
        return self._wrap(-self._a)

    def __mul__(self, other):
        if isinstance(other, Qobj):
            if self._a.shape[1] != other._a.shape[0]:
                raise TypeError("Incompatible quantum object shapes for *.")
            out = self._a @ other._a
            if out.shape == (1, 1):
                return complex(out[0, 0])
            return Qobj(out, dims=[self.dims[0], other.dims[1]], copy=False)
        if isinstance(other, numbers.Number):
            return self._wrap(self._a * other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, numbers.Number):
            return self._wrap(self._a * other)
        return NotImplemented

    __matmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, numbers.Number):
            return self._wrap(self._a / other)
        return NotImplemented

    def __pow__(self, n):
        if not self.isoper:
            raise TypeError("Only operators can be raised to a power.")
        return self._wrap(np.linalg.matrix_power(self._a, int(n)))

    def __eq__(self, other):
        if isinstance(other, Qobj):
            return (
                self.dims == other.dims
                and self._a.shape == other._a.shape
                and bool(np.allclose(self._a, other._a, atol=1e-12, rtol=0))
            )
        if isinstance(other, numbers.Number):
            # ``H != 0`` type checks in basic.py:69,75
            return bool(other == 0 and not np.any(self._a)) if other == 0 else False
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __repr__(self):
        return f"Qobj(dims={self.dims}, shape={self._a.shape}, type={self.type})\n{self._a}"

    # -- unary operations ----------------------------------------------------------------
    def dag(self):
        return Qobj(self._a.conj().T, dims=[self.dims[1], self.dims[0]], copy=False)

    def conj(self):
        return self._wrap(self._a.conj())

    def trans(self):
        return Qobj(self._a.T.copy(), dims=[self.dims[1], self.dims[0]], copy=False)

    def tr(self):
        t = np.trace(self._a)
        return float(t.real) if abs(t.imag) < 1e-15 or self.isherm else complex(t)

    def diag(self):
        d = np.diag(self._a)
        return d.real.copy() if self.isherm else d.copy()

    def norm(self):
        if self.isket or self.isbra:
            return float(np.linalg.norm(self._a))
        return float(np.sum(np.linalg.svd(self._a, compute_uv=False)))  # trace norm

    def unit(self):
        return self / self.norm()

    def expm(self):
        if not self.isoper:
            raise TypeError("expm requires an operator.")
        return self._wrap(sla.expm(self._a))

    def sqrtm(self):
        if self.isherm:
            w, v = np.linalg.eigh(self._a)
            return self._wrap((v * np.sqrt(np.clip(w, 0, None).astype(complex))) @ v.conj().T)
        return self._wrap(sla.sqrtm(self._a))

    def overlap(self, other):
        return complex(np.vdot(self._a.ravel(), other._a.ravel()))

    def matrix_element(self, bra, ket):
        return complex((bra.dag()._a if bra.isket else bra._a) @ self._a @ ket._a)

    def eigenstates(self):
        w, v = np.linalg.eigh(self._a) if self.isherm else np.linalg.eig(self._a)
        kets = [Qobj(v[:, i], dims=[self.dims[0], [1] * len(self.dims[0])]) for i in range(len(w))]
        return w, kets

    def ptrace(self, sel):
        """Partial trace keeping the subsystems in ``sel`` (in ascending order, like QuTiP)."""
        if isinstance(sel, numbers.Integral):
            sel = [int(sel)]
        sel = sorted(int(s) for s in sel)
        sub = list(self.dims[0])
        n = len(sub)
        rho = self._a if self.isoper else self._a @ self._a.conj().T
        rho = rho.reshape(sub + sub)
        keep_dims = [sub[i] for i in sel]
        for ax in sorted((i for i in range(n) if i not in sel), reverse=True):
            rho = np.trace(rho, axis1=ax, axis2=ax + rho.ndim // 2)
        d = int(np.prod(keep_dims)) if keep_dims else 1
        return Qobj(rho.reshape(d, d), dims=[keep_dims, keep_dims], copy=False)


# -- constructors -------------------------------------------------------------------------
def qeye(n) -> Qobj:
    return Qobj(np.eye(int(n)), dims=[[int(n)], [int(n)]], copy=False)


identity = qeye


def destroy(n) -> Qobj:
    n = int(n)
    return Qobj(np.diag(np.sqrt(np.arange(1, n, dtype=float)), k=1), dims=[[n], [n]], copy=False)


def create(n) -> Qobj:
    return destroy(n).dag()


def num(n) -> Qobj:
    n = int(n)
    return Qobj(np.diag(np.arange(n, dtype=float)), dims=[[n], [n]], copy=False)


def basis(n, k=0) -> Qobj:
    n, k = int(n), int(k)
    if not 0 <= k < n:
        raise ValueError("basis index out of range")
    v = np.zeros((n, 1), dtype=complex)
    v[k, 0] = 1.0
    return Qobj(v, dims=[[n], [1]], copy=False)


fock = basis


def ket2dm(q: Qobj) -> Qobj:
    if q.isket:
        return Qobj(q._a @ q._a.conj().T, dims=[q.dims[0], q.dims[0]], copy=False)
    if q.isbra:
        return Qobj(q._a.conj().T @ q._a, dims=[q.dims[1], q.dims[1]], copy=False)
    raise TypeError("ket2dm expects a ket or bra.")


def fock_dm(n, k=0) -> Qobj:
    return ket2dm(basis(n, k))


def tensor(*args) -> Qobj:
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    if not args:
        raise TypeError("tensor needs at least one argument.")
    arr = reduce(np.kron, (q._a for q in args))
    rows: List[int] = []
    cols: List[int] = []
    for q in args:
        rows += q.dims[0]
        cols += q.dims[1]
    return Qobj(arr, dims=[rows, cols], copy=False)


def displace(n, alpha) -> Qobj:
    """Truncated displacement operator ``exp(alpha a^dag - conj(alpha) a)`` (qutip.displace)."""
    a = destroy(n)._a
    return Qobj(sla.expm(alpha * a.conj().T - np.conj(alpha) * a), dims=[[int(n)], [int(n)]], copy=False)


def coherent(n, alpha) -> Qobj:
    return displace(n, alpha) * basis(n, 0)


def sigmax() -> Qobj:
    return Qobj([[0, 1], [1, 0]])


def sigmay() -> Qobj:
    return Qobj([[0, -1j], [1j, 0]])


def sigmaz() -> Qobj:
    return Qobj([[1, 0], [0, -1]])


def isket(q) -> bool:
    return isinstance(q, Qobj) and q.isket


def isoper(q) -> bool:
    return isinstance(q, Qobj) and q.isoper


# -- measures -----------------------------------------------------------------------------
def _expect_one(op: Qobj, state: Qobj):
    if state.isket:
        v = np.vdot(state._a.ravel(), op._a @ state._a.ravel())
    elif state.isbra:
        s = state._a.conj().ravel()
        v = np.vdot(s, op._a @ s)
    else:
        v = np.sum(op._a * state._a.T)
    return float(v.real) if op.isherm else complex(v)


def expect(op, state):
    """``qutip.expect``: scalar for one state, ndarray for a list of states."""
    if isinstance(op, (list, tuple)):
        return [expect(o, state) for o in op]
    if isinstance(state, (list, tuple, np.ndarray)):
        vals = [_expect_one(op, s) for s in state]
        return np.array(vals, dtype=float if op.isherm else complex)
    return _expect_one(op, state)


def _as_dm_array(q: Qobj) -> np.ndarray:
    return ket2dm(q)._a if (q.isket or q.isbra) else q._a


def fidelity(a: Qobj, b: Qobj) -> float:
    """``qutip.fidelity`` (the *square root* convention: callers square it, test_calibration.py:30)."""
    if a.isket and b.isket:
        return float(abs(np.vdot(a._a.ravel(), b._a.ravel())))
    if a.isket or b.isket:
        ket, rho = (a, b) if a.isket else (b, a)
        v = ket._a.ravel()
        return float(np.sqrt(max(np.vdot(v, rho._a @ v).real, 0.0)))
    ra, rb = _as_dm_array(a), _as_dm_array(b)
    w, v = np.linalg.eigh(0.5 * (ra + ra.conj().T))
    sq = (v * np.sqrt(np.clip(w, 0, None))) @ v.conj().T
    ev = np.linalg.eigvalsh(sq @ rb @ sq)
    return float(np.sum(np.sqrt(np.clip(ev.real, 0, None))))


def tracedist(a: Qobj, b: Qobj, sparse=False, tol=0) -> float:
    d = _as_dm_array(a) - _as_dm_array(b)
    return 0.5 * float(np.sum(np.abs(np.linalg.eigvalsh(0.5 * (d + d.conj().T)))))
