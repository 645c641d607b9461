"""``mesolve`` / ``sesolve`` / ``propagator`` front end with the QuTiP-4.x call signature.

This is the host half of the drop-in boundary: ``CompiledPulseSequence.run`` calls
``mesolve(H, init_state, tlist, c_ops, e_ops, options=...)`` exactly where the reference calls
``qutip.mesolve`` (``sequencing/sequencing/basic.py:504-512``), and gets back an object with the
``qutip.solver.Result`` attributes the reference and its tests use (``.states``, ``.expect``,
``.times``, ``.num_expect``, ``.num_collapse``, ``.solver``).  Underneath, list-format ``H`` and
``c_ops`` are lowered to dense operators + coefficient tables and integrated by the CUDA engine
through the C ABI - one trajectory, or, inside ``deferred_solves()``, every solve issued by a
sweep loop as ONE batched launch.

Semantics restated from QuTiP 4.x: ket + no collapse operators -> Schroedinger
evolution with per-output renormalisation; otherwise Lindblad evolution of a density matrix;
``[C, c]`` collapse terms contribute with rate ``spline(|c|^2)``; expectation values are real
when the operator and the initial state are Hermitian; ``store_states`` is forced on when there
are no ``e_ops``.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .. import qobj as qt
from . import _lib
from .engine import BatchOutput, BatchProblem, Engine

ODE_ERROR = (
    "ODE integration error: Try to increase the allowed number of substeps by increasing "
    "the nsteps parameter in the Options class."
)


class Options:
    """``qutip.Options`` (4.x defaults).  ``method``/``order`` are accepted and ignored: the
    engine's stepper is an embedded Runge-Kutta 5(4) pair controlled by the same atol/rtol."""

    def __init__(self, atol=1e-8, rtol=1e-6, method="adams", order=12, nsteps=1000, first_step=0,
                 max_step=0, min_step=0, store_states=False, store_final_state=False,
                 normalize_output=True, rhs_reuse=False, **ignored):
        self.atol = atol
        self.rtol = rtol
        self.method = method
        self.order = order
        self.nsteps = nsteps
        self.first_step = first_step
        self.max_step = max_step
        self.min_step = min_step
        self.store_states = store_states
        self.store_final_state = store_final_state
        self.normalize_output = normalize_output
        self.rhs_reuse = rhs_reuse
        self.kernel_method = _lib.METHOD_AUTO  # engine-specific: force a kernel (tests)
        self.order_switch = True               # engine-specific: allow the RK4(3) pair on capped steps
        for key, val in ignored.items():
            setattr(self, key, val)

    def _key(self):
        return (self.atol, self.rtol, self.nsteps, self.first_step, self.max_step, self.min_step,
                bool(self.normalize_output), int(self.kernel_method), bool(getattr(self, 'order_switch', True)))


def rhs_clear():
    """No-op: the engine keeps no cross-call RHS cache (basic.py:485 calls this before every run)."""


class Result:
    """Attribute-compatible with ``qutip.solver.Result``."""

    def __init__(self):
        self.solver = None
        self.times = None
        self.states: List[qt.Qobj] = []
        self.expect: List[np.ndarray] = []
        self.num_expect = 0
        self.num_collapse = 0
        self.ntraj = None
        self.final_state = None
        self.stats = None
        self._pending = False

    def __repr__(self):
        lines = ["Result object with " + str(self.solver) + " data.", "-" * 40]
        if self.states:
            lines.append("states = True")
        if self.expect:
            lines.append(f"expect = True\nnum_expect = {self.num_expect}, num_collapse = {self.num_collapse}")
        return "\n".join(lines)


class QobjEvo:
    """Carrier for list-format ``H`` plus the grid its array coefficients are sampled on
    (``qutip.QobjEvo(H, tlist=times)``, used for ``only_final_state`` at basic.py:500)."""

    def __init__(self, Q_object, args=None, tlist=None, state0=None, e_ops=None):
        self.spec = Q_object if isinstance(Q_object, (list, tuple)) else [Q_object]
        self.tlist = None if tlist is None else np.asarray(tlist, dtype=float)


# --------------------------------------------------------------------------------------------
# lowering
# --------------------------------------------------------------------------------------------
@dataclass
class _Lowered:
    """One trajectory in engine form."""

    H0: np.ndarray
    Hk: List[np.ndarray]
    coeff: List[np.ndarray]
    Lc: List[np.ndarray]
    Ltd: List[np.ndarray]
    rate: List[np.ndarray]
    state0: np.ndarray
    is_ket: bool
    E: List[np.ndarray]
    e_real: List[bool]
    t0: float
    dt: float
    n_t: int
    out_idx: np.ndarray
    tlist: np.ndarray
    dims: list
    options: Options
    store_states: bool
    state_herm: bool = True
    e_herm: Optional[List[bool]] = None

    src: tuple = ()      # the Qobj operators this problem was lowered from (static H terms, drive terms, collapse terms, e_ops)

    def signature(self, digest=None):
        """Grouping key of a deferred solve: shapes, grid, options and the CONTENT of every operator.  ``digest`` (a
        per-block cache, ``_Deferred._digest``) hashes each source operator once per object instead of once per solve -
        B sequences of one system share their static operators as objects (``System.frozen``)."""
        if digest is not None and self.src:
            content = tuple(digest(q) for q in self.src)
        else:
            ops = (self.H0, *self.Hk, *self.Lc, *self.Ltd, *self.E)
            content = hash(b"".join(np.ascontiguousarray(o).tobytes() for o in ops))
        return (
            self.H0.shape[0], self.is_ket, self.n_t, self.t0, self.dt, self.out_idx.tobytes(),
            len(self.Hk), len(self.Lc), len(self.Ltd), len(self.E), tuple(self.e_real),
            self.store_states, self.options._key(), content,
        )


def _is_td(term):
    return isinstance(term, (list, tuple)) and len(term) == 2 and isinstance(term[0], qt.Qobj)


def _uniform_grid(grid, need_uniform=True):
    """(t0, dt, n_t) of a uniformly spaced grid.  A constant problem (no array coefficients) may come with any
    increasing ``tlist`` (qutip.mesolve accepts that): it is then placed on the finest uniform grid that holds
    every output time, or ``ValueError`` if there is none within 2**20 points."""
    grid = np.asarray(grid, dtype=float)
    if grid.size < 2:
        return float(grid[0]) if grid.size else 0.0, 1.0, int(grid.size)
    steps = np.diff(grid)
    if np.allclose(steps, steps[0], rtol=1e-9, atol=1e-12):
        return float(grid[0]), float(steps[0]), int(grid.size)
    if need_uniform:
        raise ValueError("array coefficients need a uniformly spaced time grid")
    if np.any(steps <= 0):
        raise ValueError("tlist must be strictly increasing")
    smin = float(steps.min())
    for k in range(1, 65):          # the smallest spacing, or a simple fraction of it, as the grid step
        dt = smin / k
        pos = (grid - grid[0]) / dt
        if pos[-1] < (1 << 20) and np.all(np.abs(pos - np.rint(pos)) <= 1e-7):
            return float(grid[0]), dt, int(np.rint(pos[-1])) + 1
    raise ValueError("tlist is not commensurate with a uniform grid of <= 2**20 points")


def _lower(H, rho0, tlist, c_ops, e_ops, options) -> _Lowered:
    if isinstance(H, QobjEvo):
        grid = H.tlist if H.tlist is not None else np.asarray(tlist, dtype=float)
        H = H.spec
    else:
        grid = np.asarray(tlist, dtype=float)
        if isinstance(H, qt.Qobj):
            H = [H]
    tlist = np.asarray(tlist, dtype=float)
    has_arrays = any(_is_td(term) for term in H) or any(_is_td(term) for term in (c_ops or []))
    t0, dt, n_t = _uniform_grid(grid, need_uniform=has_arrays)
    if not isinstance(rho0, qt.Qobj):
        raise TypeError("initial state must be a Qobj")
    c_ops = list(c_ops or [])
    if isinstance(e_ops, qt.Qobj):
        e_ops = [e_ops]
    e_ops = list(e_ops or [])
    if any(callable(e) for e in e_ops):
        raise TypeError("callback e_ops are not supported by the batched engine")

    N = rho0.shape[0]
    H0 = np.zeros((N, N), dtype=complex)
    Hk, coeff = [], []
    src_static, src_td = [], []
    for term in H:
        if _is_td(term):
            src_td.append(term[0])
            c = np.asarray(term[1])
            if c.shape != (n_t,):
                raise ValueError(f"coefficient array of length {c.shape} does not match the {n_t}-point time grid")
            op = term[0]._a      # (read only from here on: stacked into the batch arrays by _to_batch)
            if np.iscomplexobj(c) and np.any(c.imag):
                # spline is linear: c H = Re(c) H + Im(c) (iH)
                Hk += [op, 1j * op]
                coeff += [np.ascontiguousarray(c.real, dtype=float), np.ascontiguousarray(c.imag, dtype=float)]
            else:
                Hk.append(op)
                coeff.append(np.ascontiguousarray(c.real, dtype=float))
        elif isinstance(term, qt.Qobj):
            H0 = H0 + term._a
            src_static.append(term)
        elif isinstance(term, (int, float, complex)) and term == 0:
            continue
        else:
            raise TypeError(f"unsupported Hamiltonian term of type {type(term)} (string/function coefficients are not used by this path)")
    Lc, Ltd, rate = [], [], []
    for term in c_ops:
        if _is_td(term):
            c = np.asarray(term[1])
            if c.shape != (n_t,):
                raise ValueError("collapse coefficient array does not match the time grid")
            Ltd.append(term[0]._a)
            rate.append(np.ascontiguousarray((np.conj(c) * c).real, dtype=float))
        elif isinstance(term, qt.Qobj):
            Lc.append(term._a)
        else:
            raise TypeError(f"unsupported collapse operator of type {type(term)}")

    is_ket = rho0.isket and not (Lc or Ltd)
    if is_ket:
        state0 = rho0.full().reshape(-1)
        dims = rho0.dims
    else:
        dm = qt.ket2dm(rho0) if rho0.isket else rho0
        state0 = dm.full()
        dims = dm.dims
    state_herm = True if is_ket else bool(np.array_equal(state0, state0.conj().T) or np.allclose(state0, state0.conj().T, atol=1e-12))
    E = [e._a for e in e_ops]
    e_real = [bool(e.isherm and state_herm) for e in e_ops]

    # outputs must sit on coefficient-grid samples (true for every call the reference makes:
    # tlist == times, or [times.min(), times.max()] for only_final_state)
    pos = (tlist - t0) / dt
    out_idx = np.rint(pos).astype(np.int64)
    if np.any(np.abs(pos - out_idx) > 1e-6) or np.any(out_idx < 0) or (n_t and np.any(out_idx > max(n_t - 1, 0))):
        if Hk or Ltd:
            raise ValueError("output times must coincide with samples of the coefficient grid")
    if np.any(np.diff(out_idx) <= 0) and out_idx.size > 1:
        raise ValueError("tlist must be strictly increasing")
    store_states = bool(options.store_states) or not e_ops
    # (static and time-dependent terms are marked apart: the same operators split differently are different problems)
    src = (*src_static, None, *src_td, None, *[t for t in c_ops if isinstance(t, qt.Qobj)], None,
           *[t[0] for t in c_ops if _is_td(t)], None, *e_ops)
    return _Lowered(H0, Hk, coeff, Lc, Ltd, rate, state0, is_ket, E, e_real, t0, dt, n_t,
                    out_idx.astype(np.int32), tlist, dims, options, store_states, state_herm,
                    [bool(e.isherm) for e in e_ops], src)


def _stack(ops, N):
    return np.stack(ops).astype(complex) if ops else np.zeros((0, N, N), dtype=complex)


def _to_batch(items: Sequence[_Lowered]) -> BatchProblem:
    first = items[0]
    N = first.H0.shape[0]
    opt = first.options
    n_t = first.n_t
    coeff = np.stack([np.stack(it.coeff) if it.coeff else np.zeros((0, n_t)) for it in items])
    rate = np.stack([np.stack(it.rate) if it.rate else np.zeros((0, n_t)) for it in items])
    return BatchProblem(
        H0=first.H0, Hk=_stack(first.Hk, N), coeff=coeff, state0=np.stack([it.state0 for it in items]),
        t0=first.t0, dt=first.dt, out_idx=first.out_idx, Lc=_stack(first.Lc, N), Ltd=_stack(first.Ltd, N),
        rate=rate, E=_stack(first.E, N), is_ket=first.is_ket, normalize=bool(opt.normalize_output),
        store_states=first.store_states, expect_complex=not all(first.e_real),
        atol=opt.atol, rtol=opt.rtol, max_step=opt.max_step or 0.0, first_step=opt.first_step or 0.0,
        min_step=opt.min_step or 0.0, nsteps=int(opt.nsteps), method=int(opt.kernel_method),
        order_switch=bool(getattr(opt, 'order_switch', True)),
    )


def _fill(result: Result, it: _Lowered, out: BatchOutput, b: int):
    if out.status[b] != _lib.STATUS_OK:
        raise Exception(ODE_ERROR + f" (engine status {int(out.status[b])} for trajectory {b})")
    result.solver = "sesolve" if it.is_ket else "mesolve"
    result.times = it.tlist
    result.num_expect = len(it.E)
    result.num_collapse = len(it.Lc) + len(it.Ltd)
    result.expect = []
    for m, real in enumerate(it.e_real):
        row = out.expect[b, m]
        result.expect.append(np.array(row.real if real else row, dtype=float if real else complex))
    st_dims = it.dims
    if it.store_states and out.states is not None:
        result.states = [qt.Qobj(s, dims=st_dims) for s in out.states[b]]
    else:
        result.states = []
    result.final_state = qt.Qobj(out.final_state[b], dims=st_dims)
    result.stats = {"accepted": int(out.stats[b, 0]), "rejected": int(out.stats[b, 1]),
                    "rhs_evals": int(out.stats[b, 2]), "kernel_ms": out.kernel_ms, "method": out.method}
    result._pending = False


# --------------------------------------------------------------------------------------------
# deferred batching: the sweep loops of calibration.py become one launch
# --------------------------------------------------------------------------------------------
class _Deferral:
    def __init__(self):
        self.items: List[_Lowered] = []
        self.results: List[Result] = []
        self.combos = []

    def add(self, it: _Lowered) -> Result:
        res = Result()
        res._pending = True
        self.items.append(it)
        self.results.append(res)
        return res

    def add_split(self, part_p: _Lowered, part_q: _Lowered) -> Result:
        """A non-Hermitian initial operator: two Hermitian trajectories, recombined when the batch has run."""
        res = Result()
        res._pending = True
        self.combos.append((res, self.add(part_p), self.add(part_q)))
        return res

    def _digest(self, q):
        """Content hash of a source operator, computed once per object inside this block (the cache keeps the array
        alive, so an id is never reused)."""
        if q is None:
            return None
        cache = self.__dict__.setdefault("_digests", {})
        a = q._a
        hit = cache.get(id(a))
        if hit is None:
            hit = (a, hash(np.ascontiguousarray(a).tobytes()))
            cache[id(a)] = hit
        return hit[1]

    def groups(self):
        groups = {}
        for i, it in enumerate(self.items):
            groups.setdefault(it.signature(self._digest), []).append(i)
        self.__dict__.pop("_digests", None)
        return list(groups.values())

    def batch_problems(self):
        """The collected solves as engine-level ``BatchProblem``s (one per compatible group)."""
        return [_to_batch([self.items[i] for i in idxs]) for idxs in self.groups()]

    def flush(self):
        for idxs in self.groups():
            items = [self.items[i] for i in idxs]
            out = _solve_maybe_sharded(_to_batch(items))
            for b, i in enumerate(idxs):
                _fill(self.results[i], self.items[i], out, b)
        for res, rp, rq in self.combos:
            _combine_parts(res, rp, rq)
        self.items, self.results, self.combos = [], [], []


def _solve_maybe_sharded(bp: BatchProblem) -> BatchOutput:
    """Under ``torch.distributed`` (one process per GPU) a deferred sweep is sharded across the
    ranks and gathered, so the same sweep script scales by launching it with torchrun."""
    try:
        import torch.distributed as dist

        active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        active = False
    if active and bp.B > 1:
        from ..distributed import solve_sharded

        return solve_sharded(bp, engine=Engine.get())
    return Engine.get().solve(bp)


_active: List[_Deferral] = []


@contextlib.contextmanager
def deferred_solves(execute=True):
    """Collect every ``mesolve``/``sesolve`` issued inside the block and run them as batched
    engine calls (grouped by shared operators / grid / options) when the block exits.  The
    ``Result`` objects returned inside the block are filled in at that point.  With
    ``execute=False`` nothing is run: ``d.batch_problems()`` gives the lowered batches."""
    d = _Deferral()
    _active.append(d)
    try:
        yield d
    finally:
        _active.remove(d)
    if execute:
        d.flush()


def _hermitian_parts(it: _Lowered):
    """Every density-matrix kernel integrates a HERMITIAN operator (the right-hand side is evaluated as
    ``A + A^dag`` and only one triangle is kept).  The Lindblad map is complex-linear, so a non-Hermitian
    initial operator X is evolved as its two Hermitian parts, X = P + iQ with P = (X + X^dag)/2 and
    Q = (X - X^dag)/(2i), and recombined: rho(t) = P(t) + i Q(t), Tr(E rho) = Tr(E P) + i Tr(E Q)."""
    import dataclasses

    X = it.state0
    P = 0.5 * (X + X.conj().T)
    Q = -0.5j * (X - X.conj().T)
    herm = list(it.e_herm) if it.e_herm is not None else [False] * len(it.E)
    return (dataclasses.replace(it, state0=P, e_real=herm, state_herm=True),
            dataclasses.replace(it, state0=Q, e_real=herm, state_herm=True))


def _combine_parts(res: Result, rp: Result, rq: Result):
    res.solver, res.times = rp.solver, rp.times
    res.num_expect, res.num_collapse = rp.num_expect, rp.num_collapse
    res.expect = [np.asarray(a, dtype=complex) + 1j * np.asarray(b, dtype=complex) for a, b in zip(rp.expect, rq.expect)]
    res.states = [qt.Qobj(a.full() + 1j * b.full(), dims=a.dims) for a, b in zip(rp.states, rq.states)]
    res.final_state = qt.Qobj(rp.final_state.full() + 1j * rq.final_state.full(), dims=rp.final_state.dims)
    res.stats = rp.stats
    res._pending = False


def mesolve(H, rho0, tlist, c_ops=None, e_ops=None, args=None, options=None, progress_bar=None, _safe_mode=True):
    """Drop-in for ``qutip.mesolve`` on the input forms the reference produces."""
    options = options or Options()
    it = _lower(H, rho0, tlist, c_ops, e_ops, options)
    if not it.is_ket and not it.state_herm:
        parts = _hermitian_parts(it)
        if _active:
            return _active[-1].add_split(*parts)
        out = Engine.get().solve(_to_batch(list(parts)))
        rp, rq, res = Result(), Result(), Result()
        _fill(rp, parts[0], out, 0)
        _fill(rq, parts[1], out, 1)
        _combine_parts(res, rp, rq)
        return res
    if _active:
        return _active[-1].add(it)
    out = Engine.get().solve(_to_batch([it]))
    res = Result()
    _fill(res, it, out, 0)
    return res


def sesolve(H, psi0, tlist, e_ops=None, args=None, options=None, progress_bar=None, _safe_mode=True):
    """Drop-in for ``qutip.sesolve`` (ket evolution)."""
    return mesolve(H, psi0, tlist, [], e_ops, args=args, options=options, progress_bar=progress_bar)


def solve_batch(bp: BatchProblem, host_abi: bool = False) -> BatchOutput:
    """Array-level public API: integrate a whole ``BatchProblem`` (host buffers in, host buffers out)."""
    eng = Engine.get()
    return eng.solve_host_abi(bp) if host_abi else eng.solve(bp)


def propagator(H, t, c_op_list=None, args=None, options=None, unitary_mode="batch", parallel=False,
               progress_bar=None, _safe_mode=True):
    """Drop-in for ``qutip.propagator`` on the forms ``basic.py:560-567`` produces.

    A propagator is a batched solve over basis states: the N basis kets (no collapse operators,
    ``U[:, n] = psi_n(t)``) or the N^2 basis operators ``|i><j|`` (with collapse operators, giving
    the column-stacked superoperator) - one engine launch either way.
    """
    options = options or Options()
    c_ops = list(c_op_list or [])
    tlist = np.atleast_1d(np.asarray(t, dtype=float))
    spec = H.spec if isinstance(H, QobjEvo) else ([H] if isinstance(H, qt.Qobj) else H)
    first = spec[0][0] if _is_td(spec[0]) else spec[0]
    N = first.shape[0]
    dims = first.dims
    opt = Options(**{k: getattr(options, k) for k in ("atol", "rtol", "nsteps", "first_step", "max_step", "min_step")})
    opt.store_states = True
    # QuTiP's unitary "batch" mode renormalises every column at every output (sesolve with
    # oper_evo and Options.normalize_output), which is what makes propagator()[-1]*psi agree with
    # run() to 5e-9 (test_sequencing.py:158); the superoperator path is never normalised.
    opt.normalize_output = bool(options.normalize_output) and not c_ops
    opt.kernel_method = options.kernel_method
    if not c_ops:
        kets = [qt.Qobj(np.eye(N)[:, n], dims=[dims[0], [1] * len(dims[0])]) for n in range(N)]
        items = [_lower(H, k, tlist, [], [], opt) for k in kets]
        out = Engine.get().solve(_to_batch(items))
        if np.any(out.status != _lib.STATUS_OK):
            raise Exception(ODE_ERROR)
        # out.states: [N(batch n), n_out, N] -> U_t[:, n]
        props = [qt.Qobj(out.states[:, k, :].T.copy(), dims=dims) for k in range(len(tlist))]
    else:
        # Hermitian operator basis (the kernels integrate Hermitian operators only): D_i = |i><i|,
        # S_ij = |i><j| + |j><i|, T_ij = -i (|i><j| - |j><i|) for i < j, so |i><j| = (S_ij + i T_ij) / 2 and
        # |j><i| = (S_ij - i T_ij) / 2.  N^2 trajectories, one launch; columns recombined below.
        basis, where = [], {}
        for i in range(N):
            m = np.zeros((N, N), dtype=complex)
            m[i, i] = 1.0
            where[("D", i)] = len(basis)
            basis.append(qt.Qobj(m, dims=dims))
        for i in range(N):
            for j in range(i + 1, N):
                m = np.zeros((N, N), dtype=complex)
                m[i, j] = m[j, i] = 1.0
                where[("S", i, j)] = len(basis)
                basis.append(qt.Qobj(m, dims=dims))
                m = np.zeros((N, N), dtype=complex)
                m[i, j], m[j, i] = -1j, 1j
                where[("T", i, j)] = len(basis)
                basis.append(qt.Qobj(m, dims=dims))
        items = [_lower(H, b, tlist, c_ops, [], opt) for b in basis]
        out = Engine.get().solve(_to_batch(items))
        if np.any(out.status != _lib.STATUS_OK):
            raise Exception(ODE_ERROR)
        sdims = [dims, dims]
        props = []
        for k in range(len(tlist)):
            sup = np.zeros((N * N, N * N), dtype=complex)
            for j in range(N):
                for i in range(N):      # column-stacking order: column index = i + j*N holds vec_F(evolved |i><j|)
                    if i == j:
                        ev = out.states[where[("D", i)], k]
                    elif i < j:
                        ev = 0.5 * (out.states[where[("S", i, j)], k] + 1j * out.states[where[("T", i, j)], k])
                    else:
                        ev = 0.5 * (out.states[where[("S", j, i)], k] - 1j * out.states[where[("T", j, i)], k])
                    sup[:, i + j * N] = ev.ravel(order="F")
            props.append(qt.Qobj(sup, dims=[[N * N], [N * N]]))
            props[-1].superdims = sdims
    if len(tlist) == 1:
        return props[0]
    return props
