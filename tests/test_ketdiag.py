"""The ket path's operators-by-diagonals kernel (``csrc/kernel_ket.cuh``, variant "ketdiag": what AUTO runs for ket
inputs without collapse operators when the Hamiltonian terms are sums of a few diagonals - ``qutip.sesolve`` semantics,
reached from ``basic.py:504``) against the dense FP64-FMA ket kernel: same stepper, same controller, same outputs."""
import os

import numpy as np
import pytest

from helpers import random_problem
from sequencing_b200.solver import _lib

pytestmark = pytest.mark.gpu


def kron_ket_problem(dims, driven, B=3, n_t=12, seed=0, hscale=0.3, **kw):
    """Product space ``dims``: Kerr / cross-Kerr / exchange terms in H0, ladder-operator drives (x and y quadrature)
    on the modes ``driven`` - the reference's Transmon / Cavity operators at random strengths - on random kets."""
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
    if len(dims) > 1:
        H0 = H0 + 0.02 * rng.normal() * (a[0] @ a[1].conj().T + a[0].conj().T @ a[1])
    Hk = []
    for m in driven:
        Hk += [a[m] + a[m].conj().T, 1j * (a[m].conj().T - a[m])]
    bp = random_problem(N=N, B=B, n_t=n_t, n_ch=len(Hk), n_c=0, n_e=2, seed=seed, hscale=hscale, ket=True, **kw)
    bp.H0 = np.ascontiguousarray(H0.astype(complex))
    bp.Hk = np.ascontiguousarray(np.stack(Hk).astype(complex))
    bp.E = np.ascontiguousarray(np.stack([n[0], a[-1] + a[-1].conj().T]).astype(complex))
    return bp


def _pair(engine, mk):
    a, g = mk(), mk()
    a.method, g.method = _lib.METHOD_DIAG, _lib.METHOD_GENERIC      # (AUTO keeps the dense kernel below ~4 terms per row of N)
    oa, og = engine.solve(a), engine.solve(g)
    assert oa.method == "diag" and oa.variant.startswith("ketdiag"), (oa.method, oa.variant)
    assert og.method == "generic"
    return oa, og


@pytest.mark.parametrize("dims,driven", [((3, 3), (0,)), ((3, 3, 3), (0, 2)), ((3,) * 5, (0, 3, 1)), ((2, 9), (0, 1)), ((40,), (0,)), ((3, 4, 25), (0,))])
def test_ketdiag_matches_dense_ket_kernel(gpu_engine, dims, driven):
    """1 .. 2 elements per thread (N = 9 .. 300), sesolve renormalisation on and off, stored states, capped and free steps."""
    for kw in ({"store_states": True}, {"store_states": True, "normalize": False}, {"max_step": 0.0, "hscale": 1.5}):
        mk = lambda: kron_ket_problem(dims, driven, seed=len(dims) + sum(dims), **kw)
        oa, og = _pair(gpu_engine, mk)
        assert np.all(oa.status == 0) and np.all(og.status == 0)
        assert np.max(np.abs(oa.final_state - og.final_state)) < 1e-11
        assert np.max(np.abs(oa.expect - og.expect)) < 1e-11
        if oa.states is not None:
            assert np.max(np.abs(oa.states - og.states)) < 1e-11
        np.testing.assert_array_equal(oa.stats[:, :3], og.stats[:, :3])
        if kw.get("normalize", True):
            assert np.max(np.abs(np.linalg.norm(oa.final_state, axis=1) - 1)) < 1e-12


def test_ketdiag_complex_expect_scaled_table_and_failures(gpu_engine):
    def mk():
        bp = kron_ket_problem((3, 3, 4), (0, 1), B=4, seed=7)
        rng = np.random.default_rng(3)
        bp.E = bp.E + 0.3 * (rng.normal(size=bp.E.shape) + 1j * rng.normal(size=bp.E.shape))      # dense, non-Hermitian
        bp.expect_complex = True
        bp.coeff = bp.coeff[:1]
        bp.coeff_scale = rng.uniform(0.2, 3.0, size=(4, bp.Hk.shape[0]))
        return bp
    oa, og = _pair(gpu_engine, mk)
    assert np.iscomplexobj(oa.expect) and np.max(np.abs(oa.expect - og.expect)) < 1e-11

    def too_many_steps():
        bp = kron_ket_problem((3, 3), (0,), seed=5, hscale=30.0)
        bp.max_step, bp.nsteps = 0.0, 3
        return bp
    oa, og = _pair(gpu_engine, too_many_steps)
    assert np.all(og.status == _lib.STATUS_MAX_STEPS)
    np.testing.assert_array_equal(oa.status, og.status)
    np.testing.assert_array_equal(oa.stats[:, :3], og.stats[:, :3])


def test_dense_kets_keep_the_generic_kernel(gpu_engine):
    auto = gpu_engine.solve(kron_ket_problem((3,) * 5, (0, 3, 1)))
    assert auto.method == "diag" and auto.variant.startswith("ketdiag")      # what AUTO runs on a five-qutrit register
    out = gpu_engine.solve(random_problem(N=20, B=2, n_t=6, n_ch=2, n_c=0, n_e=2, seed=5, ket=True))
    assert out.method == "generic"
    bp = random_problem(N=20, B=2, n_t=6, n_ch=2, n_c=0, n_e=2, seed=5, ket=True)
    bp.method = _lib.METHOD_DIAG            # forced: dense operators have 2N-1 diagonals each -> too many terms, loud failure
    with pytest.raises(RuntimeError):
        gpu_engine.solve(bp)
    os.environ["SEQ_KET_DIAG"] = "0"
    try:
        assert gpu_engine.solve(kron_ket_problem((3, 3), (0,))).method == "generic"
    finally:
        del os.environ["SEQ_KET_DIAG"]
