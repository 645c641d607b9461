"""CPU oracle: a restatement of the QuTiP-4.x ``mesolve`` / ``sesolve`` algorithm.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / the timed CPU baseline.

What it restates
----------------
The reference hands its time evolution to a third-party dependency that is *not* vendored under
``/root/reference``: ``qutip>=4.5`` (``requirements.txt:2``, ``setup.py:24``; 4.x-only API, the
docs were executed on 4.5.2).  The single call site is
``sequencing/sequencing/basic.py:504-512`` (``qutip.mesolve(H, init_state, tlist, c_ops, e_ops,
options=...)``), with options set at ``basic.py:480-485`` (``max_step = dt``,
``store_states = True``).  QuTiP is not installed in the build container and cannot be fetched,
so this file restates its published 4.x algorithm:

* ``liouvillian`` / ``lindblad_dissipator`` in the column-stacking convention
  (``spre(A) = I (x) A``, ``spost(A) = A^T (x) I``)            [qutip/superoperator.py]
* array coefficients -> natural cubic spline on the uniform ``tlist``
  (``_prep_cubic_spline``, ``_spline_*_cte_second``)          [qutip/cy/inter.pyx]
* RHS = ``L0.v + sum_k c_k(t) (L_k.v)``, one CSR SpMV per term [qutip/cy/cqobjevo.pyx mul_vec]
* ``scipy.integrate.ode('zvode', method='adams', order=12, atol=1e-8, rtol=1e-6, nsteps=1000,
  first_step=0, min_step=0, max_step=opt.max_step)`` stepped with ``integrate(r.t + dt)`` per
  output sample                                                [qutip/mesolve.py _generic_ode_solve]
* ``expect_rho_vec`` = Tr(E rho), real part when E and rho0 are Hermitian
* ket + no collapse ops -> ``sesolve``: RHS = ``(-i H(t)) psi``, renormalise at every output and
  re-initialise the integrator when ``abs(norm-1) > 1e-12``     [qutip/sesolve.py]
* time-dependent collapse operator ``[C, c]``: dissipator coefficient is the spline of
  ``conj(c)*c`` sampled at the knots (QobjEvo product of array coefficients).

scipy's ZVODE is the same Fortran integrator QuTiP 4.x drives, so the integrator *binary
family* is shared with the reference.

PARITY PINNING STATUS: **parity unpinned at the numeric-trace level** - the reference holds no
golden vectors or stored traces for this path and QuTiP itself cannot be run here.  The oracle is
pinned instead against every known-answer test the reference's own suite holds for the path
(``tests/test_oracle_reference_kats.py``: pi-pulse infidelity < 1e-10
(test_calibration.py:31), T1 fit within 0.1 ns incl. time-dependent collapse coefficients
(test_sequencing.py:195-257), displacement infidelity < 1e-7 (test_calibration.py:79), ...),
against scipy's independent ``CubicSpline(bc_type='natural')`` for the interpolation, and
against a DOP853 rtol=1e-13 "truth" run for error attribution.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.sparse as sp
from scipy.integrate import ode, solve_ivp
from scipy.linalg import solve_banded

ArrayOrPair = Union[np.ndarray, Tuple[np.ndarray, np.ndarray], list]


# --------------------------------------------------------------------------------------------
# qutip.Options (4.x defaults; qutip/solver.py)
# --------------------------------------------------------------------------------------------
@dataclass
class Options:
    atol: float = 1e-8
    rtol: float = 1e-6
    method: str = "adams"
    order: int = 12
    nsteps: int = 1000
    first_step: float = 0.0
    min_step: float = 0.0
    max_step: float = 0.0
    store_states: bool = False
    store_final_state: bool = False
    normalize_output: bool = True


@dataclass
class OracleResult:
    solver: str
    times: np.ndarray
    expect: List[np.ndarray] = field(default_factory=list)
    states: List[np.ndarray] = field(default_factory=list)
    final_state: Optional[np.ndarray] = None
    nfev: int = 0


# --------------------------------------------------------------------------------------------
# natural cubic spline (qutip/cy/inter.pyx)
# --------------------------------------------------------------------------------------------
def prep_cubic_spline(y: np.ndarray, dt: float) -> np.ndarray:
    """Second derivatives ``M`` of the natural cubic spline through ``y`` on a uniform grid.

    Solves ``M[i-1]/2 + 2 M[i] + M[i+1]/2 = 3 (y[i+1] - 2 y[i] + y[i-1]) / dt^2`` with
    ``M[0] = M[n-1] = 0`` (``_prep_cubic_spline``, equal-spacing branch, via
    ``scipy.linalg.solve_banded`` as QuTiP does).
    """
    y = np.asarray(y)
    n = y.shape[0]
    M = np.zeros(n, dtype=y.dtype if np.iscomplexobj(y) else float)
    if n < 3:
        return M
    rhs = 3.0 * (y[2:] - 2.0 * y[1:-1] + y[:-2]) / (dt * dt)
    ab = np.zeros((3, n - 2))
    ab[0, 1:] = 0.5
    ab[1, :] = 2.0
    ab[2, :-1] = 0.5
    M[1:-1] = solve_banded((1, 1), ab, rhs)
    return M


def spline_eval(x: float, t0: float, dt: float, y: np.ndarray, M: np.ndarray):
    """``_spline_{float,complex}_cte_second``: clamp outside the grid, cubic inside."""
    n = y.shape[0]
    if x < t0:
        return y[0]
    if x > t0 + (n - 1) * dt:
        return y[n - 1]
    p = int((x - t0) / dt)
    if p >= n - 1:
        return y[n - 1]
    tb = (x - (t0 + p * dt)) / dt
    te = 1.0 - tb
    c = dt * dt / 6.0
    Mb = M[p] * c
    Me = M[p + 1] * c
    return te * (Mb * te * te + (y[p] - Mb)) + tb * (Me * tb * tb + (y[p + 1] - Me))


# --------------------------------------------------------------------------------------------
# superoperators, column-stacking convention (qutip/superoperator.py)
# --------------------------------------------------------------------------------------------
def _csr(a) -> sp.csr_matrix:
    m = a.tocsr().astype(complex) if sp.issparse(a) else sp.csr_matrix(np.asarray(a, dtype=complex))
    m.eliminate_zeros()
    return m


def spre(A):
    A = _csr(A)
    return sp.kron(sp.identity(A.shape[0], format="csr"), A, format="csr")


def spost(A):
    A = _csr(A)
    return sp.kron(A.T, sp.identity(A.shape[0], format="csr"), format="csr")


def liouvillian_h(H):
    """``-i (spre(H) - spost(H))``."""
    return (-1j * (spre(H) - spost(H))).tocsr()


def lindblad_dissipator(a):
    a = _csr(a)
    ad_a = (a.conj().T @ a).tocsr()
    return (spre(a) @ spost(a.conj().T) - 0.5 * spre(ad_a) - 0.5 * spost(ad_a)).tocsr()


def _is_pair(term) -> bool:
    return isinstance(term, (list, tuple)) and len(term) == 2 and not np.isscalar(term[0])


def _isherm(a: np.ndarray) -> bool:
    a = np.asarray(a)
    return a.ndim == 2 and a.shape[0] == a.shape[1] and np.allclose(a, a.conj().T, atol=1e-12)


class _TdOperator:
    """``cte + sum_k spline_k(t) * op_k`` acting on a vector (CQobjEvoTd.mul_vec)."""

    def __init__(self, size: int, t0: float, dt: float):
        self.cte = sp.csr_matrix((size, size), dtype=complex)
        self.ops: List[sp.csr_matrix] = []
        self.ys: List[np.ndarray] = []
        self.Ms: List[np.ndarray] = []
        self.t0 = t0
        self.dt = dt
        self.nfev = 0

    def add_const(self, m):
        self.cte = (self.cte + m).tocsr()

    def add_td(self, m, y):
        y = np.asarray(y)
        y = y.astype(complex) if np.iscomplexobj(y) else y.astype(float)
        self.ops.append(m.tocsr())
        self.ys.append(y)
        self.Ms.append(prep_cubic_spline(y, self.dt))

    def coeffs(self, t):
        return [spline_eval(t, self.t0, self.dt, y, M) for y, M in zip(self.ys, self.Ms)]

    def mul_vec(self, t, v):
        self.nfev += 1
        out = self.cte @ v
        for op, y, M in zip(self.ops, self.ys, self.Ms):
            out = out + spline_eval(t, self.t0, self.dt, y, M) * (op @ v)
        return out


def _grid(coeff_tlist):
    coeff_tlist = np.asarray(coeff_tlist, dtype=float)
    t0 = float(coeff_tlist[0])
    dt = float(coeff_tlist[1] - coeff_tlist[0]) if coeff_tlist.size > 1 else 1.0
    return t0, dt


def build_liouvillian(H: Sequence[ArrayOrPair], c_ops: Sequence[ArrayOrPair], coeff_tlist) -> _TdOperator:
    """``_mesolve_QobjEvo``: L(t) = liouvillian(H(t)) + sum_j lindblad_dissipator(C_j(t))."""
    t0, dt = _grid(coeff_tlist)
    first = H[0][0] if _is_pair(H[0]) else H[0]
    n = np.asarray(first).shape[0]
    L = _TdOperator(n * n, t0, dt)
    for term in H:
        if _is_pair(term):
            L.add_td(liouvillian_h(term[0]), term[1])
        else:
            L.add_const(liouvillian_h(term))
    for term in c_ops:
        if _is_pair(term):
            c = np.asarray(term[1])
            L.add_td(lindblad_dissipator(term[0]), (np.conj(c) * c).real)
        else:
            L.add_const(lindblad_dissipator(term))
    return L


def build_schrodinger(H: Sequence[ArrayOrPair], coeff_tlist) -> _TdOperator:
    """``_sesolve_QobjEvo``: ``-i H(t)``."""
    t0, dt = _grid(coeff_tlist)
    first = H[0][0] if _is_pair(H[0]) else H[0]
    n = np.asarray(first).shape[0]
    S = _TdOperator(n, t0, dt)
    for term in H:
        if _is_pair(term):
            S.add_td(-1j * _csr(term[0]), term[1])
        else:
            S.add_const(-1j * _csr(term))
    return S


def _make_zvode(func, opt: Options):
    r = ode(func)
    r.set_integrator(
        "zvode",
        method=opt.method,
        order=opt.order,
        atol=opt.atol,
        rtol=opt.rtol,
        nsteps=opt.nsteps,
        first_step=opt.first_step,
        min_step=opt.min_step,
        max_step=opt.max_step,
    )
    return r


_ODE_ERROR = (
    "ODE integration error: Try to increase the allowed number of substeps by increasing "
    "the nsteps parameter in the Options class."
)


def mesolve(
    H: Sequence[ArrayOrPair],
    rho0: np.ndarray,
    tlist,
    c_ops: Sequence[ArrayOrPair] = (),
    e_ops: Sequence[np.ndarray] = (),
    options: Optional[Options] = None,
    coeff_tlist=None,
) -> OracleResult:
    """QuTiP-4.x ``mesolve`` for list-format ``H`` with array coefficients.

    ``H``/``c_ops`` entries are dense ``[N,N]`` arrays (constant) or ``(array, coeff[n_t])``
    pairs.  ``coeff_tlist`` is the grid the coefficient arrays live on (``QobjEvo(H,
    tlist=times)`` in ``basic.py:500``); it defaults to ``tlist``.
    """
    opt = options or Options()
    tlist = np.asarray(tlist, dtype=float)
    if coeff_tlist is None:
        coeff_tlist = tlist
    rho0 = np.asarray(rho0, dtype=complex)
    isket = rho0.ndim == 1 or (rho0.ndim == 2 and rho0.shape[1] == 1)
    if isket and len(c_ops) == 0:
        return sesolve(H, rho0, tlist, e_ops, opt, coeff_tlist)
    if isket:
        psi = rho0.reshape(-1)
        rho0 = np.outer(psi, psi.conj())
    n = rho0.shape[0]
    L = build_liouvillian(H, c_ops, coeff_tlist)
    r = _make_zvode(L.mul_vec, opt)
    r.set_initial_value(rho0.ravel("F"), tlist[0])

    rho_herm = _isherm(rho0)
    e_data = [np.asarray(e, dtype=complex) for e in e_ops]
    e_herm = [_isherm(e) and rho_herm for e in e_data]
    res = OracleResult(solver="mesolve", times=tlist)
    res.expect = [np.zeros(len(tlist), dtype=float if h else complex) for h in e_herm]
    dts = np.diff(tlist)
    for i in range(len(tlist)):
        if not r.successful():
            raise Exception(_ODE_ERROR)
        rho_t = r.y.reshape((n, n), order="F")
        if opt.store_states:
            res.states.append(rho_t.copy())
        for m, e in enumerate(e_data):
            val = np.sum(e * rho_t.T)  # Tr(E rho)
            res.expect[m][i] = val.real if e_herm[m] else val
        if i < len(tlist) - 1:
            r.integrate(r.t + dts[i])
    res.final_state = r.y.reshape((n, n), order="F").copy()
    res.nfev = L.nfev
    return res


def sesolve(
    H: Sequence[ArrayOrPair],
    psi0: np.ndarray,
    tlist,
    e_ops: Sequence[np.ndarray] = (),
    options: Optional[Options] = None,
    coeff_tlist=None,
) -> OracleResult:
    """QuTiP-4.x ``sesolve`` (ket evolution), including ``normalize_output`` restarts."""
    opt = options or Options()
    tlist = np.asarray(tlist, dtype=float)
    if coeff_tlist is None:
        coeff_tlist = tlist
    psi0 = np.asarray(psi0, dtype=complex).reshape(-1)
    S = build_schrodinger(H, coeff_tlist)
    r = _make_zvode(S.mul_vec, opt)
    r.set_initial_value(psi0, tlist[0])
    e_data = [np.asarray(e, dtype=complex) for e in e_ops]
    e_herm = [_isherm(e) for e in e_data]
    res = OracleResult(solver="sesolve", times=tlist)
    res.expect = [np.zeros(len(tlist), dtype=float if h else complex) for h in e_herm]
    dts = np.diff(tlist)
    cdata = psi0.copy()
    for i in range(len(tlist)):
        if not r.successful():
            raise Exception(_ODE_ERROR)
        cdata = np.array(r.y, dtype=complex)
        if opt.normalize_output:
            nrm = np.linalg.norm(cdata)
            cdata = cdata / nrm
            if abs(1.0 / nrm - 1.0) > 1e-12:
                r.set_initial_value(cdata, r.t)
        if opt.store_states:
            res.states.append(cdata.copy())
        for m, e in enumerate(e_data):
            val = np.vdot(cdata, e @ cdata)
            res.expect[m][i] = val.real if e_herm[m] else val
        if i < len(tlist) - 1:
            r.integrate(r.t + dts[i])
    res.final_state = cdata.copy()
    res.nfev = S.nfev
    return res


def truth_solve(
    H: Sequence[ArrayOrPair],
    rho0: np.ndarray,
    tlist,
    c_ops: Sequence[ArrayOrPair] = (),
    e_ops: Sequence[np.ndarray] = (),
    coeff_tlist=None,
    rtol: float = 1e-13,
    atol: float = 1e-15,
    max_step: float = np.inf,
) -> OracleResult:
    """Converged solution of the same ODE (DOP853) for error attribution.

    Not the reference's algorithm - used only to say *whose* error a parity miss is.
    """
    tlist = np.asarray(tlist, dtype=float)
    if coeff_tlist is None:
        coeff_tlist = tlist
    rho0 = np.asarray(rho0, dtype=complex)
    isket = rho0.ndim == 1 or (rho0.ndim == 2 and rho0.shape[1] == 1)
    ket_mode = isket and len(c_ops) == 0
    if ket_mode:
        op = build_schrodinger(H, coeff_tlist)
        y0 = rho0.reshape(-1)
    else:
        if isket:
            psi = rho0.reshape(-1)
            rho0 = np.outer(psi, psi.conj())
        op = build_liouvillian(H, c_ops, coeff_tlist)
        y0 = rho0.ravel("F")
    # integrate knot-to-knot so the C2-only spline joints never sit inside a step
    t0g, dtg = _grid(coeff_tlist)
    sol_y = [y0]
    y = y0
    for a, b in zip(tlist[:-1], tlist[1:]):
        knots = t0g + dtg * np.arange(np.ceil((a - t0g) / dtg - 1e-9), np.floor((b - t0g) / dtg + 1e-9) + 1)
        pts = np.unique(np.concatenate([[a], knots[(knots > a) & (knots < b)], [b]]))
        for p, q in zip(pts[:-1], pts[1:]):
            s = solve_ivp(op.mul_vec, (p, q), y, method="DOP853", rtol=rtol, atol=atol, max_step=max_step)
            y = s.y[:, -1]
        sol_y.append(y)
    n = int(round(np.sqrt(y0.size))) if not ket_mode else y0.size
    res = OracleResult(solver="truth", times=tlist)
    e_data = [np.asarray(e, dtype=complex) for e in e_ops]
    res.expect = [np.zeros(len(tlist), dtype=complex) for _ in e_data]
    for i, yv in enumerate(sol_y):
        if ket_mode:
            st = yv / np.linalg.norm(yv)
            for m, e in enumerate(e_data):
                res.expect[m][i] = np.vdot(st, e @ st)
        else:
            st = yv.reshape((n, n), order="F")
            for m, e in enumerate(e_data):
                res.expect[m][i] = np.sum(e * st.T)
        res.states.append(st)
    res.final_state = res.states[-1]
    res.nfev = op.nfev
    return res


def trace_distance(a: np.ndarray, b: np.ndarray) -> float:
    """``0.5 * ||a - b||_1`` for density matrices; for kets, the distance of the projectors."""
    a = np.asarray(a, dtype=complex)
    b = np.asarray(b, dtype=complex)
    if a.ndim == 1 or (a.ndim == 2 and a.shape[1] == 1):
        a = np.outer(a.reshape(-1), a.reshape(-1).conj())
    if b.ndim == 1 or (b.ndim == 2 and b.shape[1] == 1):
        b = np.outer(b.reshape(-1), b.reshape(-1).conj())
    d = a - b
    d = 0.5 * (d + d.conj().T)
    return 0.5 * float(np.sum(np.abs(np.linalg.eigvalsh(d))))
