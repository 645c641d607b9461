"""Block-local diagonal kernel (``csrc/kernel_diagblk.cuh``: one tensor factor of the Hilbert space kept inside a thread,
rolled stage loop, 2 CTAs per SM) against the FP64-FMA kernel and the plain diagonal kernel: same equations, same stepper,
identical step decisions.  The reference builds exactly such operators - ``Mode.a / n`` embedded by Kronecker products
(``modes.py:125-207``, ``system.py:399-445``)."""
import os

import numpy as np
import pytest

from helpers import random_problem, row_sparse_problem
from sequencing_b200 import workloads
from sequencing_b200.solver import _lib

pytestmark = pytest.mark.gpu


def _solve(engine, bp, method):
    bp.method = method
    return engine.solve(bp)


def _same(a, g, tol=1e-12):
    assert np.all(a.status == 0) and np.all(g.status == 0)
    for k in ("final_state", "expect"):
        assert np.max(np.abs(getattr(a, k) - getattr(g, k))) < tol, k
    if a.states is not None and g.states is not None:
        assert np.max(np.abs(a.states - g.states)) < tol
    np.testing.assert_array_equal(a.stats[:, :3], g.stats[:, :3])


def kron_problem(dims, local, B=3, n_t=10, seed=0, drive_other=False, complex_l=False, td_collapse=False, store_states=True, hscale=0.3, max_step=None,
                 exchange=True):
    """A product space ``dims`` with ladder-operator drives on mode ``local`` (and optionally on another mode), Kerr /
    cross-Kerr / exchange terms in H0, decay + excitation + dephasing on every mode - the operator structure of
    ``System.H0`` / ``System.c_ops`` - at random strengths."""
    rng = np.random.default_rng(seed)
    N = int(np.prod(dims))

    def embed(op, m):
        out = np.eye(1)
        for q, d in enumerate(dims):
            out = np.kron(out, op if q == m else np.eye(d))
        return out

    a = [embed(np.diag(np.sqrt(np.arange(1, d)), 1), m) for m, d in enumerate(dims)]
    n = [x.conj().T @ x for x in a]
    H0 = sum(rng.normal() * hscale * x + rng.normal() * 0.1 * x @ x for x in n)
    for p in range(len(dims)):
        for q in range(p + 1, len(dims)):
            H0 = H0 + rng.normal() * 0.05 * n[p] @ n[q]
    if len(dims) > 1 and exchange:
        other = (local + 1) % len(dims)
        g = 0.02 * rng.normal()
        H0 = H0 + g * (a[local] @ a[other].conj().T + a[local].conj().T @ a[other])      # exchange coupling: a gathered diagonal
    Hk = [a[local] + a[local].conj().T, 1j * (a[local].conj().T - a[local])]
    if drive_other and len(dims) > 1:
        other = (local + 1) % len(dims)
        Hk += [a[other] + a[other].conj().T, 1j * (a[other].conj().T - a[other])]
    Ls = []
    for m in range(len(dims)):
        ph = np.exp(1j * rng.uniform(0, 6.28)) if complex_l else 1.0
        Ls += [np.sqrt(rng.uniform(0.002, 0.02)) * ph * a[m], np.sqrt(rng.uniform(0.0005, 0.002)) * a[m].conj().T, np.sqrt(rng.uniform(0.002, 0.01)) * n[m]]
    n_ctd = 1 if td_collapse else 0
    bp = random_problem(N=N, B=B, n_t=n_t, n_ch=len(Hk), n_c=len(Ls) - n_ctd, n_ctd=n_ctd, n_e=2, seed=seed, hscale=hscale, store_states=store_states)
    bp.H0 = np.ascontiguousarray(H0.astype(complex))
    bp.Hk = np.ascontiguousarray(np.stack(Hk).astype(complex))
    bp.Lc = np.ascontiguousarray(np.stack(Ls[:len(Ls) - n_ctd]).astype(complex))
    if n_ctd:
        bp.Ltd = np.ascontiguousarray(np.stack(Ls[len(Ls) - n_ctd:]).astype(complex))
    bp.E = np.ascontiguousarray(np.stack([n[local], n[(local + 1) % len(dims)]]).astype(complex))
    if max_step is not None:
        bp.max_step = max_step
    return bp


@pytest.mark.parametrize("dims,local", [((3, 5), 0), ((5, 3), 1), ((3, 3, 4), 1), ((2, 7), 0), ((6, 2), 1), ((3, 15), 0), ((2, 2, 6), 0), ((3, 3, 10), 0)])
def test_blk_matches_generic_and_plain_on_product_spaces(gpu_engine, dims, local):
    """Local mode first / last / in the middle of the Kronecker order (stride N/D, 1, in between), dimensions 2 and 3,
    one and several hi-blocks, every launch shape (128 x 2 CTAs/SM, 256, 512 with streamed stage derivatives), gathered
    G diagonals (exchange coupling, drive on another mode) next to the local ones, complex collapse weights, a
    time-dependent collapse rate, capped and free steps."""
    for kw in ({}, {"drive_other": True, "complex_l": True}, {"td_collapse": True, "hscale": 2.0, "max_step": 0.0}):
        mk = lambda: kron_problem(dims, local, seed=sum(dims) + local, **kw)
        blk = _solve(gpu_engine, mk(), _lib.METHOD_DIAG)
        assert blk.method == "diag" and blk.variant.startswith("diagblk"), blk.variant
        gen = _solve(gpu_engine, mk(), _lib.METHOD_GENERIC)
        # (the dense kernel's own rounding - N-term sums, Hermiticity not enforced: 3e-11 .. 1.3e-10 at N = 45 over 300 steps;
        # the plain diagonal kernel below is the sharper comparison)
        _same(blk, gen, tol=1e-9)
        os.environ["SEQ_DIAG_BLK"] = "0"
        try:
            plain = _solve(gpu_engine, mk(), _lib.METHOD_DIAG)
        finally:
            del os.environ["SEQ_DIAG_BLK"]
        assert not plain.variant.startswith("diagblk")
        _same(blk, plain, tol=1e-11 if "hscale" in kw else 1e-12)


@pytest.mark.parametrize("dims,local", [((3, 5), 0), ((3, 4), 0), ((4, 3), 1), ((2, 3, 4), 1), ((3, 3, 3), 0), ((3, 15), 0)])
def test_blk_owned_half_only_when_every_gather_stays_in_the_band(gpu_engine, dims, local):
    """Only the local mode is driven and nothing couples the modes by exchange: every gathered term is a collapse pair
    (o, o) of another mode, which moves an outer pair along its own band position - the kernel then writes the owned
    orientation of every block only (no mirror writes) and resolves the expectation operators' entries at kernel start.
    Odd and even outer dimensions (the half-way band position is owned by the lower rows only), one and several
    hi-blocks, expectation operators with off-diagonal entries, against the FP64-FMA kernel."""
    def mk():
        bp = kron_problem(dims, local, seed=3 + sum(dims), exchange=False, store_states=False, n_t=14)
        rng = np.random.default_rng(1)
        h = rng.normal(size=bp.E[0].shape) + 1j * rng.normal(size=bp.E[0].shape)
        bp.E = bp.E.copy()
        bp.E[1] = bp.E[1] + 0.1 * (h + h.conj().T) * (np.abs(np.subtract.outer(np.arange(bp.N), np.arange(bp.N))) <= 2)      # a few off-diagonals
        return bp
    blk = _solve(gpu_engine, mk(), _lib.METHOD_DIAG)
    assert blk.variant.startswith("diagblk") and "no mirror writes" in blk.variant, blk.variant
    _same(blk, _solve(gpu_engine, mk(), _lib.METHOD_GENERIC), tol=1e-10)
    st = mk()
    st.store_states = True          # stored states read the full matrix: the mirror comes back
    assert "no mirror writes" not in _solve(gpu_engine, st, _lib.METHOD_DIAG).variant


def test_blk_is_what_auto_runs_on_the_named_workloads(gpu_engine):
    """C2 (transmon(3) x cavity(15)) keeps the qubit inside a thread: 1 shared-memory gather per element instead of 6; C4
    (two transmons + bus) keeps one transmon.  Against the plain diagonal kernel: identical step counts, 1e-13."""
    for bp_mk, gath in ((lambda: workloads.c2_selective(13, 12, sigma=20), "1 gathers/element (plain diag: 6)"), (lambda: workloads.c4_two_transmons_bus(5), None)):
        auto = gpu_engine.solve(bp_mk())
        assert auto.method == "diag" and auto.variant.startswith("diagblk"), auto.variant
        if gath:
            assert gath in auto.variant, auto.variant
        os.environ["SEQ_DIAG_BLK"] = "0"
        try:
            plain = gpu_engine.solve(bp_mk())
        finally:
            del os.environ["SEQ_DIAG_BLK"]
        assert plain.method == "diag" and not plain.variant.startswith("diagblk")
        _same(auto, plain, tol=1e-13)
    small = gpu_engine.solve(workloads.c2_selective(3, 5, sigma=20))      # fewer trajectories than SMs: latency-bound, the plain kernel is faster
    assert small.method == "diag" and not small.variant.startswith("diagblk")


def test_blk_random_diagonals_take_it_only_when_a_diagonal_is_local(gpu_engine):
    """Random diagonal-structured operators (no tensor structure): a candidate (stride, dimension) is accepted only if
    the DATA has a diagonal that never leaves its block; whatever runs must equal the FP64-FMA kernel."""
    seen = set()
    for N, w_h, w_l, n_c, n_ctd in ((12, 3, 1, 2, 1), (12, 1, 1, 2, 0), (20, 5, 2, 2, 1), (24, 3, 1, 2, 0), (30, 3, 2, 3, 1), (45, 3, 1, 4, 0)):
        mk = lambda: row_sparse_problem(N=N, B=3, n_t=8, w_h=w_h, w_l=w_l, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N, hscale=0.3, store_states=True)
        a = _solve(gpu_engine, mk(), _lib.METHOD_DIAG)
        seen.add(a.variant.split(":")[0])
        _same(a, _solve(gpu_engine, mk(), _lib.METHOD_GENERIC))
    assert seen == {"diag", "diagblk"}, seen


def test_blk_failure_statuses_match_generic(gpu_engine):
    """nsteps exceeded (capped and free steps) and a non-finite state: status words and step counts of the FP64-FMA kernel."""
    def capped():
        bp = kron_problem((3, 5), 0, seed=21, hscale=0.05, store_states=False)
        bp.max_step, bp.nsteps = bp.dt / 3, 2
        return bp

    def uncapped():
        bp = kron_problem((3, 5), 0, seed=22, hscale=30.0, store_states=False)
        bp.max_step, bp.nsteps = 0.0, 3
        return bp

    def nonfinite():
        bp = kron_problem((3, 5), 0, seed=23, store_states=False)
        bp.state0 = bp.state0.copy()
        bp.state0[1, 0, 0] = np.nan
        return bp

    for mk, expect in ((capped, _lib.STATUS_MAX_STEPS), (uncapped, _lib.STATUS_MAX_STEPS), (nonfinite, None)):
        blk, gen = _solve(gpu_engine, mk(), _lib.METHOD_DIAG), _solve(gpu_engine, mk(), _lib.METHOD_GENERIC)
        assert blk.variant.startswith("diagblk")
        np.testing.assert_array_equal(blk.status, gen.status)
        if expect is not None:
            assert np.all(gen.status == expect)
            np.testing.assert_array_equal(blk.stats[:, :3], gen.stats[:, :3])
        else:
            assert gen.status[1] == _lib.STATUS_NONFINITE and gen.status[0] == 0 and gen.status[2] == 0
            assert np.max(np.abs(blk.final_state[[0, 2]] - gen.final_state[[0, 2]])) < 1e-12


def test_blk_scaled_shared_table_and_complex_expect(gpu_engine):
    """One shared coefficient table with per-trajectory scales (amplitude sweeps) and non-Hermitian e_ops."""
    def mk():
        bp = kron_problem((3, 6), 0, B=5, seed=9)
        rng = np.random.default_rng(2)
        bp.E = bp.E + 0.3 * (rng.normal(size=bp.E.shape) + 1j * rng.normal(size=bp.E.shape))
        bp.expect_complex = True
        bp.coeff = bp.coeff[:1]
        bp.coeff_scale = rng.uniform(0.2, 3.0, size=(5, bp.Hk.shape[0]))
        return bp
    blk = _solve(gpu_engine, mk(), _lib.METHOD_DIAG)
    assert blk.variant.startswith("diagblk")
    _same(blk, _solve(gpu_engine, mk(), _lib.METHOD_GENERIC))


def test_blk_without_expectation_operators_and_single_trajectory(gpu_engine):
    """n_e = 0 (only final states are asked for) and B = 1: the owned-half mode needs no expectation lists at all."""
    def mk(B):
        bp = kron_problem((3, 5), 0, B=B, seed=11, exchange=False, store_states=False)
        bp.E = np.zeros((0, bp.N, bp.N), complex)
        return bp
    for B in (1, 4):
        blk = _solve(gpu_engine, mk(B), _lib.METHOD_DIAG)
        assert blk.variant.startswith("diagblk") and "no mirror writes" in blk.variant, blk.variant
        gen = _solve(gpu_engine, mk(B), _lib.METHOD_GENERIC)
        assert blk.expect.shape[1] == 0
        assert np.max(np.abs(blk.final_state - gen.final_state)) < 1e-11
        np.testing.assert_array_equal(blk.stats[:, :3], gen.stats[:, :3])
