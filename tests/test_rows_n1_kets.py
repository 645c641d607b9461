"""Rows the round-1 verdict found untested on the product path:

* N1 - ``propagator`` (``basic.py:514-576``) with and without collapse operators, and ``mesolve`` on a
  NON-HERMITIAN initial operator (every density-matrix kernel integrates a Hermitian operator, so
  the host splits X = P + iQ into its Hermitian parts and recombines) - CPU (oracle engine: pins the split /
  recombination logic) and GPU (the real engine against the oracle's full-Liouvillian solve);
* a10 - the ket path at the reference's largest ket case, 5 qutrits (N = 243, ``test_qasm.py:463-530`` shape):
  pulsed single-qutrit rotations, engine vs oracle ``sesolve`` and vs the ideal rotation;
* ``run_sequences`` handing kets to a segment that carries collapse terms;
* constant Hamiltonians on a non-uniform ``tlist`` and ``out_idx`` validation.
"""
import numpy as np
import pytest

import sequencing_b200.qobj as qt
from helpers import oracle_inputs, oracle_options
from oracle import qutip4_oracle as orc
from sequencing_b200 import CTerm, Cavity, Operation, Sequence, System, Transmon, get_sequence, sync
from sequencing_b200.solver import Options, api


def _lossy_qubit_cavity(levels_c=4):
    q = Transmon("qubit", levels=3, kerr=-0.2, t1=2e3, t2=1.5e3)
    c = Cavity("cavity", levels=levels_c, kerr=-1e-4, t1=5e3)
    system = System("system", modes=[q, c])
    system.set_cross_kerr(q, c, -2e-3)
    return system, q, c


def _lowered(system, build, state):
    with api.deferred_solves(execute=False) as d:
        seq = get_sequence(system)
        build()
        seq.run(state, only_final_state=True)
    (bp,) = d.batch_problems()
    return bp


def _oracle_final(bp, X):
    """Final operator of the oracle's full-Liouvillian solve started from an arbitrary operator X."""
    H, _, tlist, c_ops, e_ops, grid = oracle_inputs(bp, 0)
    r = orc.mesolve(H, X, tlist, c_ops, [], oracle_options(bp, atol=1e-12, rtol=1e-10, nsteps=1000000), coeff_tlist=grid)
    return np.asarray(r.final_state)


def _check_nonhermitian_and_propagator(engine_kind):
    system, q, c = _lossy_qubit_cavity()
    N = 12
    rng = np.random.default_rng(3)
    X = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
    X /= np.linalg.norm(X)
    Xq = qt.Qobj(X, dims=[[3, 4], [3, 4]])
    opt = Options(atol=1e-12, rtol=1e-10, nsteps=100000)
    opt.max_step = system.dt
    opt.store_states = True
    seq = get_sequence(system)
    q.rotate_x(np.pi / 2)
    c.displace(0.4)
    cseq = seq.compile()
    # (1) mesolve on a non-Hermitian operator, with a non-Hermitian e_op -> complex expectation values
    E = qt.Qobj(rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)), dims=Xq.dims)
    res = cseq.run(Xq, e_ops=[E], options=opt)
    bp = _lowered(system, lambda: (q.rotate_x(np.pi / 2), c.displace(0.4)), system.fock_dm())
    ref = _oracle_final(bp, X)
    assert np.max(np.abs(res.states[-1].full() - ref)) < 2e-8
    assert np.iscomplexobj(res.expect[0])
    assert abs(res.expect[0][-1] - np.trace(E.full() @ ref)) < 2e-7        # (|E|_F ~ 17)
    assert abs(res.expect[0][0] - np.trace(E.full() @ X)) < 1e-12
    # (2) propagator with collapse operators = column-stacked superoperator: acts on vec_F(X)
    props = cseq.propagator(options=opt)
    sup = props[-1].full()
    assert sup.shape == (N * N, N * N)
    got = (sup @ X.ravel(order="F")).reshape((N, N), order="F")
    assert np.max(np.abs(got - ref)) < 2e-8
    # trace preservation of the map: sum over i of row (i + i N) is vec(I)^T
    tr_row = sum(sup[i + i * N] for i in range(N))
    assert np.max(np.abs(tr_row - np.eye(N).ravel(order="F"))) < 1e-8


def test_nonhermitian_rho0_and_superoperator_propagator_cpu(oracle_engine):
    _check_nonhermitian_and_propagator("oracle")


@pytest.mark.gpu
def test_nonhermitian_rho0_and_superoperator_propagator_gpu(gpu_engine):
    _check_nonhermitian_and_propagator("cuda")


@pytest.mark.gpu
def test_deferred_nonhermitian_split_gpu(gpu_engine):
    """The split also works inside ``deferred_solves`` (two hidden trajectories per non-Hermitian operator)."""
    system, q, c = _lossy_qubit_cavity()
    rng = np.random.default_rng(4)
    Xs = [rng.normal(size=(12, 12)) + 1j * rng.normal(size=(12, 12)) for _ in range(3)]
    Xs = [qt.Qobj(X / np.linalg.norm(X), dims=[[3, 4], [3, 4]]) for X in Xs]       # (unit Frobenius norm: absolute bounds below)
    opt = Options(atol=1e-12, rtol=1e-10, nsteps=100000)
    opt.max_step = system.dt
    res = []
    with api.deferred_solves():
        for X in Xs:
            seq = get_sequence(system)
            q.rotate_x(np.pi / 2)
            res.append(seq.run(X, e_ops=[q.n], options=opt, only_final_state=True))
    bp = _lowered(system, lambda: q.rotate_x(np.pi / 2), system.fock_dm())
    for X, r in zip(Xs, res):
        ref = _oracle_final(bp, X.full())
        assert np.max(np.abs(r.final_state.full() - ref)) < 2e-8
        assert abs(r.expect[0][-1] - np.trace(q.n.full() @ ref)) < 2e-8


@pytest.mark.gpu
def test_superoperator_propagator_n45_gpu(gpu_engine):
    """C2's system (N = 45, n_c = 4): 2025 Hermitian basis trajectories in one launch; the superoperator applied to
    random operators against the oracle's direct solves."""
    q = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
    c = Cavity("cavity", levels=15, kerr=-1e-5, t1=100e3, t2=80e3)
    system = System("system", modes=[q, c])
    system.set_cross_kerr(c, q, chi=-2e-3)
    N = 45
    opt = Options(atol=1e-12, rtol=1e-10, nsteps=100000)
    opt.max_step = system.dt
    seq = get_sequence(system)
    q.rotate_x(np.pi)
    cseq = seq.compile()
    # the assembly of CompiledPulseSequence.propagator (basic.py:540-567), outputs at the two ends only
    c_ops = []
    H, times = cseq._assemble(c_ops)
    sup = api.propagator(api.QobjEvo(H, tlist=times), [times.min(), times.max()], c_op_list=c_ops, options=opt)[-1].full()
    bp = _lowered(system, lambda: q.rotate_x(np.pi), system.fock_dm())
    rng = np.random.default_rng(5)
    for _ in range(2):
        X = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
        X /= np.linalg.norm(X)
        ref = _oracle_final(bp, X)
        got = (sup @ X.ravel(order="F")).reshape((N, N), order="F")
        assert np.max(np.abs(got - ref)) < 5e-8


def _five_qutrits():
    qs = [Transmon(f"q{i}", levels=3, kerr=-0.25) for i in range(5)]
    system = System("system", modes=qs)
    return system, qs


@pytest.mark.gpu
def test_ket_path_n243_five_qutrits(gpu_engine):
    """Kets at N = 243 (``test_qasm.py:463-530``: 5 qutrits, pulsed single-qubit gates between unitaries): two pulsed
    rotations on different qutrits from a superposition, engine (sesolve semantics: kets out, renormalised) vs the
    oracle's ``sesolve`` at matched tight tolerances and vs the ideal rotations."""
    system, qs = _five_qutrits()
    assert system.I().shape[0] == 243
    rng = np.random.default_rng(6)
    psi = np.zeros(243, complex)
    psi[0] = 1.0
    psi[rng.integers(0, 243, 6)] += 0.3 * (rng.normal(size=6) + 1j * rng.normal(size=6))
    psi /= np.linalg.norm(psi)
    init = qt.Qobj(psi.reshape(-1, 1), dims=[[3] * 5, [1] * 5])
    opt = Options(atol=1e-13, rtol=1e-11, nsteps=100000)      # matched with the oracle run below
    opt.max_step = system.dt
    opt.store_states = True

    def build():
        qs[0].rotate_y(np.pi / 2)
        qs[3].rotate_x(np.pi / 3)
        sync()
        qs[1].rotate_y(np.pi)

    seq = get_sequence(system)
    build()
    res = seq.run(init, e_ops=[qs[0].n, qs[1].n], options=opt)
    assert res.solver == "sesolve" and res.states[-1].isket
    bp = _lowered(system, build, init)
    assert bp.is_ket and bp.N == 243
    H, s0, tlist, c_ops, e_ops, grid = oracle_inputs(bp, 0)
    # the reference trace: the oracle's sesolve CONVERGED (1e-13 / 1e-11).  At 1e-11 / 1e-9 it is still 3e-6 from this
    # one - sesolve renormalises at every output and restarts ZVODE at order 1, 120 restarts here
    ref = orc.sesolve(H, psi, grid, [qs[0].n.full(), qs[1].n.full()], oracle_options(bp, atol=1e-13, rtol=1e-11, nsteps=1000000, store_states=False), coeff_tlist=grid)
    got = res.states[-1].full().reshape(-1)
    assert abs(np.linalg.norm(got) - 1) < 1e-12
    # (the oracle's own 1e-13/1e-11 and 1e-14/1e-12 runs are 1.3e-8 apart on this problem: 5e-8 is its resolution)
    assert np.linalg.norm(got - np.asarray(ref.final_state).reshape(-1)) < 5e-8
    for m in range(2):
        assert np.max(np.abs(res.expect[m] - np.real(ref.expect[m]))) < 5e-8
    # default tolerances (atol 1e-8, rtol 1e-6): neither integrator is within 1e-6 of the converged trace on 120 ns of
    # free stepping (the oracle's ZVODE run: trace distance 4e-4, expectation traces 1.7e-5) - the engine must be at least
    # as close to it as the reference integrator is
    ref_d = orc.sesolve(H, psi, grid, [qs[0].n.full(), qs[1].n.full()], oracle_options(bp, store_states=False), coeff_tlist=grid)
    res_d = get_sequence(system)
    build()
    res_d = res_d.run(init, e_ops=[qs[0].n, qs[1].n])
    fin = np.asarray(ref.final_state).reshape(-1)
    e_eng = max(np.max(np.abs(res_d.expect[m] - np.real(ref.expect[m]))) for m in range(2))
    e_orc = max(np.max(np.abs(np.real(ref_d.expect[m]) - np.real(ref.expect[m]))) for m in range(2))
    td_eng = orc.trace_distance(res_d.states[-1].full().reshape(-1), fin)
    td_orc = orc.trace_distance(np.asarray(ref_d.final_state).reshape(-1), fin)
    print(f"ket N=243 default tolerances vs converged: engine td {td_eng:.2e} expect {e_eng:.2e} | oracle td {td_orc:.2e} expect {e_orc:.2e}")
    assert e_eng <= max(1e-6, e_orc) and td_eng <= max(1e-6, td_orc)


@pytest.mark.gpu
def test_run_sequences_kets_meet_collapse_segment(gpu_engine):
    """Ket inputs, a lossless system, and ONE segment that carries a CTerm: the states turn into density matrices at
    that segment (what ``qutip.mesolve`` returns) and stay density matrices - against the per-state ``run`` loop."""
    q = Transmon("qubit", levels=3, kerr=-0.2)
    c = Cavity("cavity", levels=4, kerr=-1e-4)
    system = System("system", modes=[q, c])
    system.set_cross_kerr(q, c, -2e-3)

    def build():
        seq = Sequence(system)
        q.rotate_x(np.pi / 2)
        seq.capture()
        seq.append(q.Ry(np.pi / 3))
        op = q.rotate_y(np.pi, capture=False)
        terms = dict(op.terms)
        terms["qubit.loss"] = CTerm(q.a, coeffs=0.05)
        seq.append(Operation(op.duration, terms))
        seq.append(q.Rx(np.pi / 5))
        c.displace(0.3)
        seq.capture()
        return seq

    inits = [system.fock(qubit=0, cavity=0), (system.fock(qubit=1, cavity=0) + 1j * system.fock(qubit=0, cavity=2)).unit()]
    e_ops = [q.n, c.n]
    batch = build().run_batch(inits, e_ops=e_ops)
    for st, got in zip(inits, batch):
        ref = build().run(st, e_ops=e_ops, full_evolution=False)
        assert [s.isket for s in got.states] == [s.isket for s in ref.states]
        assert not got.states[-1].isket and got.states[1].isket
        for a, b in zip(got.states, ref.states):
            assert np.max(np.abs(a.full() - b.full())) < 1e-11
        for a, b in zip(got.expect, ref.expect):
            np.testing.assert_allclose(a, b, atol=1e-11)


@pytest.mark.gpu
def test_constant_hamiltonian_nonuniform_tlist_and_out_idx_validation(gpu_engine):
    """qutip.mesolve accepts a constant H with any increasing tlist; the C ABI rejects out-of-range / unordered out_idx."""
    from helpers import random_problem
    q = Transmon("qubit", levels=3, kerr=-0.2, t1=50.0)
    tlist = np.array([0.0, 0.5, 2.0, 2.25, 7.0])
    rho0 = qt.ket2dm((q.fock(1) + q.fock(2)).unit())
    res = api.mesolve(q.n * 0.3 + (q.a + q.ad) * 0.2, rho0, tlist, [np.sqrt(1 / 50.0) * q.a], [q.n], options=Options(atol=1e-12, rtol=1e-10, nsteps=100000))
    H = [(q.n * 0.3 + (q.a + q.ad) * 0.2).full()]
    ref = orc.mesolve(H, rho0.full(), tlist, [np.sqrt(1 / 50.0) * q.a.full()], [q.n.full()], orc.Options(atol=1e-12, rtol=1e-10, nsteps=1000000))
    assert np.max(np.abs(res.expect[0] - ref.expect[0])) < 1e-8
    with pytest.raises(ValueError):
        api.mesolve([q.n, [q.a + q.ad, np.ones(5)]], rho0, tlist, [], [q.n])      # array coefficients need a uniform grid
    for bad in ([0, 3, 2], [0, 40], [-1, 3]):
        bp = random_problem(N=3, B=2, n_t=10, n_ch=1, n_c=1, n_e=1, seed=1)
        bp.out_idx = np.array(bad, dtype=np.int32)
        with pytest.raises(RuntimeError):
            gpu_engine.solve(bp)
        with pytest.raises(RuntimeError):
            gpu_engine.solve_host_abi(bp)
