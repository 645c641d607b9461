"""Synthetic sweeps of the named benchmark shapes (BASELINE.json configs).

Each builder drives the public host API (``System`` / ``Transmon`` / ``Cavity`` /
``get_sequence`` / ``seq.run``) inside ``deferred_solves(execute=False)`` and returns the lowered
``BatchProblem``: the same operators and coefficient tables a user's sweep loop would hand to
the engine.  Everything is deterministic; anything random uses ``numpy.random.default_rng(0)``.
"""
import numpy as np

from . import qobj as qt
from .modes import Cavity, Transmon
from .sequencing import get_sequence, sync
from .solver import Options, deferred_solves
from .system import CouplingTerm, System


def _options(dt, store_states=False, **kw):
    opt = Options(**kw)
    opt.max_step = dt           # the reference default (basic.py:482)
    opt.store_states = store_states
    return opt


def _single(batches):
    if len(batches) != 1:
        raise RuntimeError(f"workload lowered to {len(batches)} batches, expected one")
    return batches[0]


def c1_rabi(B=51, levels=3, store_states=False):
    """C1: single transmon(3) Rabi amplitude sweep (``tune_rabi`` shape) with T1/T2 collapse ops.
    N=3, n_t=40, n_ch=2, n_c=2."""
    qubit = Transmon("qubit", levels=levels, kerr=-0.2, t1=20e3, t2=15e3)
    system = System("system", modes=[qubit])
    init = qubit.fock_dm(0)
    with deferred_solves(execute=False) as d:
        for amp in np.linspace(-2, 2, B):
            seq = get_sequence(system)
            with qubit.pulse_scale(amp):
                qubit.rotate_x(np.pi)
            seq.run(init, e_ops=[init], options=_options(system.dt, store_states))
    return _single(d.batch_problems())


def c2_selective(n_amp=32, n_det=32, sigma=250, cavity_levels=15, store_states=False):
    """C2: transmon(3) (x) cavity(15) dispersive system, number-selective pi pulse,
    amplitude x detuning sweep.  N=45, n_t=4*sigma, n_ch=2, n_c=4, B=n_amp*n_det."""
    qubit = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
    cavity = Cavity("cavity", levels=cavity_levels, kerr=-1e-5, t1=100e3, t2=80e3)
    system = System("system", modes=[qubit, cavity])
    chi = -2e-3
    system.set_cross_kerr(cavity, qubit, chi=chi)
    qubit.gaussian_pulse.sigma = sigma
    rng = np.random.default_rng(0)
    e_ops = [system.fock_dm(qubit=1, cavity=n) for n in range(4)]
    with deferred_solves(execute=False) as d:
        for amp in np.linspace(0.5, 1.5, n_amp):
            for k in np.linspace(0.0, 3.0, n_det):
                n0 = int(rng.integers(0, 4))
                seq = get_sequence(system)
                with qubit.pulse_scale(amp):
                    qubit.rotate_x(np.pi, detune=k * chi)
                seq.run(system.fock_dm(qubit=0, cavity=n0), e_ops=e_ops, options=_options(system.dt, store_states))
    return _single(d.batch_problems())


def cardinal_states(qubit):
    z0, z1 = qubit.fock(0), qubit.fock(1)
    kets = [z0, z1, (z0 + z1).unit(), (z0 - z1).unit(), (z0 + 1j * z1).unit(), (z0 - 1j * z1).unit()]
    return [qt.ket2dm(k) for k in kets]


def c3_drag(n_drag=4096, levels=4, family="XpY9", with_loss=True, store_states=False):
    """C3: DRAG sweep (``tune_drag`` shape) on transmon(4): ``R_x(pi);sync;R_y(pi/2)`` (XpY9) or
    ``R_y(pi);sync;R_x(pi/2)`` (YpX9) x 6 cardinal initial states.  N=4, n_t=80, B=6*n_drag."""
    kw = dict(t1=20e3, t2=15e3) if with_loss else {}
    qubit = Transmon("qubit", levels=levels, kerr=-0.2, **kw)
    system = System("system", modes=[qubit])
    pulse = qubit.gaussian_pulse
    states = cardinal_states(qubit)
    e_ops = [qubit.fock_dm(1)]
    with deferred_solves(execute=False) as d:
        for drag in np.linspace(-5, 5, n_drag):
            pulse.drag = drag
            for init in states:
                seq = get_sequence(system)
                if family == "XpY9":
                    qubit.rotate_x(np.pi); sync(); qubit.rotate_y(np.pi / 2)
                else:
                    qubit.rotate_y(np.pi); sync(); qubit.rotate_x(np.pi / 2)
                seq.run(init, e_ops=e_ops, options=_options(system.dt, store_states))
    pulse.drag = 0
    return _single(d.batch_problems())


def c4_two_transmons_bus(B=512, cavity_levels=10, store_states=False):
    """C4: two transmons(3) + shared bus cavity(10) (N=90) with cross-Kerrs and an exchange
    coupling, full collapse-operator set (thermal population 0.01 -> n_c=9), pulsed ``h``-like
    ``R_y(pi/2)`` on both qubits then ``R_y`` rotations with amplitude/phase jitter ~N(0,1e-2)."""
    common = dict(levels=3, kerr=-0.2, t1=20e3, t2=15e3, thermal_population=0.01)
    q0 = Transmon("qubit0", **common)
    q1 = Transmon("qubit1", **common)
    cav = Cavity("cavity", levels=cavity_levels, kerr=-1e-5, t1=100e3, t2=80e3, thermal_population=0.01)
    system = System("system", modes=[q0, q1, cav])
    system.set_cross_kerr(cav, q0, chi=-2e-3)
    system.set_cross_kerr(cav, q1, chi=-1.5e-3)
    system.set_cross_kerr(q0, q1, chi=-1e-4)
    key = frozenset([q0.name, cav.name])
    system.coupling_terms[key].append(CouplingTerm([(q0, "a"), (cav, "ad")], strength=2 * np.pi * 1e-3, add_hc=True))
    rng = np.random.default_rng(0)
    init = system.fock_dm()
    e_ops = [q0.fock_dm(1), q1.fock_dm(1)]
    with deferred_solves(execute=False) as d:
        for _ in range(B):
            da, dp = rng.normal(0, 1e-2, size=2)
            seq = get_sequence(system)
            with q0.pulse_scale(1 + da), q1.pulse_scale(1 + da):
                q0.rotate(np.pi / 2, np.pi / 2 + dp)
                q1.rotate(np.pi / 2, np.pi / 2 - dp)
                sync()
                q0.rotate(np.pi, dp)
                cav.displace(0.5)
            seq.run(init, e_ops=e_ops, options=_options(system.dt, store_states))
    return _single(d.batch_problems())


def c5_transmon_two_cavities(cavity_levels=25, n_ops=100, sigma=10, store_states=False):
    """C5: transmon(3) (x) cavity(L) (x) cavity(L) (N = 3 L^2 = 1875 at L=25), one long sequence of
    ``n_ops`` alternating cavity displacements and qubit rotations (n_t = 4*sigma*n_ops), decay on all
    three modes + transmon dephasing (n_c = 6: decay and dephasing on every mode).  B = 1: the single
    large system whose rho is row-sharded across the GPUs of the box."""
    qubit = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
    cav_a = Cavity("cavity_a", levels=cavity_levels, kerr=-1e-5, t1=100e3, t2=80e3)
    cav_b = Cavity("cavity_b", levels=cavity_levels, kerr=-1e-5, t1=120e3, t2=90e3)
    system = System("system", modes=[qubit, cav_a, cav_b])
    system.set_cross_kerr(cav_a, qubit, chi=-2e-3)
    system.set_cross_kerr(cav_b, qubit, chi=-1.5e-3)
    system.set_cross_kerr(cav_a, cav_b, chi=-1e-5)
    for mode in (qubit, cav_a, cav_b):
        mode.gaussian_pulse.sigma = sigma
    rng = np.random.default_rng(0)
    init = system.fock_dm()
    e_ops = [qubit.fock_dm(1), cav_a.n, cav_b.n]
    with deferred_solves(execute=False) as d:
        seq = get_sequence(system)
        for i in range(n_ops):
            which = i % 3
            if which == 0:
                cav_a.displace(0.3 * np.exp(1j * rng.uniform(0, 2 * np.pi)))
            elif which == 1:
                qubit.rotate(rng.uniform(0.2, 1.0) * np.pi, rng.uniform(0, 2 * np.pi))
            else:
                cav_b.displace(0.3 * np.exp(1j * rng.uniform(0, 2 * np.pi)))
            sync()
        seq.run(init, e_ops=e_ops, options=_options(system.dt, store_states))
    return _single(d.batch_problems())
