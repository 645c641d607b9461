"""Pin the CPU oracle (oracle/qutip4_oracle.py) against independent references: scipy's natural
cubic spline, dense Lindblad algebra, closed-form solutions and a converged DOP853 run."""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline

from oracle import qutip4_oracle as orc


def test_spline_matches_scipy_natural_cubic_spline():
    rng = np.random.default_rng(0)
    for n_t, dt, t0 in ((3, 1.0, 0.0), (7, 0.75, 2.0), (40, 1.0, 0.0), (200, 2.0, -5.0)):
        t = t0 + dt * np.arange(n_t)
        y = rng.normal(size=n_t)
        cs = CubicSpline(t, y, bc_type="natural")
        M = orc.prep_cubic_spline(y, dt)
        np.testing.assert_allclose(M, cs(t, 2), atol=1e-11)
        for x in np.concatenate([rng.uniform(t[0], t[-1], 50), t]):
            assert abs(orc.spline_eval(x, t0, dt, y, M) - cs(x)) < 1e-12
        # clamped outside the grid (qutip returns the end samples, not an extrapolation)
        assert orc.spline_eval(t[0] - 3.0, t0, dt, y, M) == y[0]
        assert orc.spline_eval(t[-1] + 3.0, t0, dt, y, M) == y[-1]


def test_spline_complex_and_short_tables():
    y = np.array([1 + 2j, 0.5 - 1j, 0.25 + 0.5j, -1j])
    M = orc.prep_cubic_spline(y, 0.5)
    Mr, Mi = orc.prep_cubic_spline(y.real, 0.5), orc.prep_cubic_spline(y.imag, 0.5)
    np.testing.assert_allclose(M, Mr + 1j * Mi, atol=1e-14)
    assert np.all(orc.prep_cubic_spline(np.array([1.0, 2.0]), 1.0) == 0)


def test_liouvillian_matches_dense_lindblad_form():
    rng = np.random.default_rng(1)
    N = 5
    H = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)); H = H + H.conj().T
    Ls = [rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)) for _ in range(3)]
    rho = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    L = orc.build_liouvillian([H], Ls, np.arange(4.0))
    got = (L.cte @ rho.ravel("F")).reshape((N, N), order="F")
    want = -1j * (H @ rho - rho @ H)
    for c in Ls:
        cd = c.conj().T
        want += c @ rho @ cd - 0.5 * (cd @ c @ rho + rho @ cd @ c)
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_constant_drive_rabi_closed_form():
    """H = (Omega/2) sigma_x as an array coefficient: P_1(t) = sin^2(Omega t / 2)."""
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    omega = 0.2
    t = np.arange(0, 41.0)
    psi0 = np.array([1, 0], dtype=complex)
    P1 = np.diag([0.0, 1.0]).astype(complex)
    opt = orc.Options(max_step=1.0, atol=1e-12, rtol=1e-10, nsteps=10000)
    r = orc.mesolve([np.zeros((2, 2)), (sx, 0.5 * omega * np.ones_like(t))], psi0, t, [], [P1], opt)
    assert r.solver == "sesolve"
    np.testing.assert_allclose(r.expect[0], np.sin(omega * t / 2) ** 2, atol=2e-9)
    rho0 = np.outer(psi0, psi0.conj())
    r2 = orc.mesolve([np.zeros((2, 2)), (sx, 0.5 * omega * np.ones_like(t))], rho0, t, [], [P1], opt)
    assert r2.solver == "mesolve"
    np.testing.assert_allclose(r2.expect[0], np.sin(omega * t / 2) ** 2, atol=2e-9)


def test_t1_decay_closed_form_static_and_time_dependent_collapse():
    """Excited-state population decays as exp(-t/T1); a CTerm with coefficient sqrt(g) adds its
    rate (dissipator coefficient = spline(|c|^2), test_sequencing.py:195-257)."""
    a = np.array([[0, 1], [0, 0]], dtype=complex)
    P1 = np.diag([0.0, 1.0]).astype(complex)
    rho0 = P1.copy()
    t = np.arange(0, 200.0)
    T1 = 100.0
    opt = orc.Options(max_step=1.0, atol=1e-12, rtol=1e-10, nsteps=10000)
    r = orc.mesolve([np.zeros((2, 2))], rho0, t, [np.sqrt(1 / T1) * a], [P1], opt)
    np.testing.assert_allclose(r.expect[0], np.exp(-t / T1), atol=1e-9)
    extra = 1 / 50.0 - 1 / T1
    r = orc.mesolve([np.zeros((2, 2))], rho0, t, [np.sqrt(1 / T1) * a, (a, np.sqrt(extra) * np.ones_like(t))], [P1], opt)
    np.testing.assert_allclose(r.expect[0], np.exp(-t / 50.0), atol=1e-9)


def test_default_tolerance_run_is_close_to_truth_and_tight_run_is_closer():
    rng = np.random.default_rng(3)
    N, n_t = 4, 40
    h = lambda: (lambda m: 0.15 * (m + m.conj().T))(rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)))
    t = np.arange(n_t, dtype=float)
    H = [h(), (h(), np.sin(0.3 * t)), (h(), np.cos(0.2 * t))]
    c_ops = [0.1 * (rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)))]
    rho0 = np.diag([1.0, 0, 0, 0]).astype(complex)
    E = [np.diag(np.arange(N, dtype=float)).astype(complex)]
    truth = orc.truth_solve(H, rho0, t, c_ops, E)
    d = orc.mesolve(H, rho0, t, c_ops, E, orc.Options(max_step=1.0))
    tight = orc.mesolve(H, rho0, t, c_ops, E, orc.Options(max_step=1.0, atol=1e-10, rtol=1e-8))
    td_d = orc.trace_distance(d.final_state, truth.final_state)
    td_t = orc.trace_distance(tight.final_state, truth.final_state)
    assert td_d < 2e-5 and td_t < 2e-7 and td_t < td_d
    assert abs(np.trace(d.final_state) - 1) < 1e-6


def test_sesolve_normalises_outputs_and_only_final_state_grid():
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    t = np.arange(0, 30.0)
    c = 0.1 * np.exp(-((t - 15) / 6.0) ** 2)
    psi0 = np.array([1, 0], dtype=complex)
    full = orc.sesolve([np.zeros((2, 2)), (sx, c)], psi0, t, [], orc.Options(max_step=1.0, store_states=True))
    for s in full.states:
        assert abs(np.linalg.norm(s) - 1) < 1e-12
    ends = orc.sesolve([np.zeros((2, 2)), (sx, c)], psi0, [t[0], t[-1]], [], orc.Options(max_step=1.0, store_states=True), coeff_tlist=t)
    assert len(ends.states) == 2
    assert orc.trace_distance(ends.final_state, full.final_state) < 1e-5


def test_integration_failure_raises_like_qutip():
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    t = np.arange(0, 5.0)
    with pytest.raises(Exception, match="ODE integration error"):
        orc.mesolve([500.0 * sx], np.array([1, 0], dtype=complex), t, [], [], orc.Options(nsteps=5))
