"""The reference's own known-answer tests for the hot path, run through the
host API mirror twice (fixture ``kat_engine``): with the CPU oracle standing in for the GPU
(``-m "not gpu"``: pins the host mirror - System/Mode/pulses/channel compiler/calibration sweeps -
and the oracle; these are the only result pins the reference holds for this path, it has no golden
vectors) and with the REAL CUDA engine behind the same calls (``-m gpu``: the reference's own
acceptance thresholds - pi-pulse infidelity < 1e-10, fitted T1 within 0.1 ns with time-dependent
collapse coefficients over 10 040 samples, propagator vs run < 5e-9, displacement infidelity
< 1e-7 - on the product path)."""
import numpy as np
import pytest

import sequencing_b200.qobj as qt
from sequencing_b200 import (CTerm, Cavity, Operation, Sequence, System, Transmon, capture_operation, delay,
                             get_sequence, sync)
from sequencing_b200.calibration import tune_displacement, tune_drag, tune_rabi
from sequencing_b200.sequencing import CompiledPulseSequence, PulseSequence, SyncOperation
from sequencing_b200.solver import Result


def test_rabi_two_levels(kat_engine):
    """test_calibration.py:16-31: amp converges < 1e-5, pi-pulse infidelity < 1e-10."""
    qubit = Transmon("qubit", levels=2)
    system = System("system", modes=[qubit])
    system.dt = 0.75
    for _ in range(5):
        _, old_amp, new_amp = tune_rabi(system, qubit.fock(0), progbar=False, plot=False)
    assert abs(old_amp - new_amp) < 1e-5
    # every sweep was ONE batched engine call of 51 trajectories
    assert set(kat_engine.batch_sizes) == {51}
    init = qubit.fock(0)
    seq = get_sequence(system)
    qubit.rotate_x(np.pi)
    result = seq.run(init)
    fidelity = qt.fidelity(result.states[-1], qubit.Rx(np.pi) * init) ** 2
    assert abs(1 - fidelity) < 1e-10


def test_drag(kat_engine):
    """test_calibration.py:41-61: 3-level transmon, kerr -0.2 GHz."""
    qubit = Transmon("qubit", levels=3, kerr=-200e-3)
    qubit.gaussian_pulse.sigma = 10
    system = System("system", modes=[qubit])
    for _ in range(5):
        _, old_amp, new_amp = tune_rabi(system, qubit.fock(0), plot=False, verify=False, progbar=False)
    assert abs(old_amp - new_amp) < 1e-7
    _, old_drag, new_drag = tune_drag(system, qubit.fock(0), update=True, progbar=False, plot=False)
    assert abs(new_drag) > 1e-3
    assert kat_engine.batch_sizes[-1] == 42   # 21 drag values x 2 sequences in one launch
    init = qubit.fock(0)
    seq = get_sequence(system)
    qubit.rotate_x(np.pi)
    result = seq.run(init)
    fidelity = qt.fidelity(result.states[-1], qubit.Rx(np.pi) * init) ** 2
    assert abs(1 - fidelity) < 1e-5


def test_displacement(kat_engine):
    """test_calibration.py:65-79: cavity(12) displacement calibration, infidelity < 1e-7."""
    cavity = Cavity("cavity", levels=12)
    system = System("system", modes=[cavity])
    for _ in range(3):
        _, old_amp, new_amp = tune_displacement(system, cavity.fock(0), progbar=False, plot=False)
    assert abs(old_amp - new_amp) < 1e-7
    init = cavity.fock(0)
    seq = get_sequence(system)
    cavity.displace(1 + 2j)
    result = seq.run(init)
    fidelity = qt.fidelity(result.states[-1], cavity.D(1 + 2j) * init) ** 2
    assert abs(1 - fidelity) < 1e-7


def test_pulse_sequence_and_propagator(kat_engine):
    """test_sequencing.py:122-193: sizes, Result type, propagator vs run < 5e-9."""
    qubit = Transmon("qubit", levels=2)
    pulse_len = qubit.gaussian_pulse.sigma * qubit.gaussian_pulse.chop
    system = System("system", modes=[qubit])
    seq = PulseSequence(system=system)
    for i in range(3):
        seq.append(qubit.rotate_x(np.pi, capture=False))
        if i < 2:
            seq.append(SyncOperation())
    assert isinstance(seq.compile(), CompiledPulseSequence)
    assert seq.compile().hc.times.size == 3 * pulse_len
    prop = seq.propagator()
    result = seq.run(qubit.fock(0))
    assert isinstance(result, Result) and result.solver == "sesolve"
    assert all(isinstance(p, qt.Qobj) for p in prop) and len(prop) == 3 * pulse_len
    fid = qt.fidelity(prop[-1] * qubit.fock(0), result.states[-1]) ** 2
    assert abs(1 - fid) < 5e-9

    cseq = CompiledPulseSequence(system=system)
    cseq.add_operation(qubit.rotate_x(np.pi, capture=False))
    cseq.sync()
    cseq.add_operation(qubit.rotate_x(np.pi, capture=False))
    assert cseq.hc.times.size == 2 * pulse_len
    assert isinstance(cseq.run(qubit.fock(0)), Result)


def test_dynamic_collapse_operators(kat_engine):
    """test_sequencing.py:195-257: fitted T1 within 0.1 ns, static and time-dependent CTerm."""
    qubit = Transmon("qubit", levels=2)
    qubit.t1 = 1000
    system = System("system", modes=[qubit])

    @capture_operation
    def lossy_pi_pulse(qubit, pulsed_t1):
        coeff = np.sqrt(1 / pulsed_t1 - qubit.Gamma_down)
        op = qubit.rotate_x(np.pi, capture=False)
        terms = op.terms
        terms[f"{qubit.name}.Gamma_down"] = CTerm(qubit.a, coeffs=coeff)
        return Operation(op.duration, terms)

    @capture_operation
    def lossy_delay(qubit, length, pulsed_t1):
        coeff = np.sqrt(1 / pulsed_t1 - qubit.Gamma_down)
        return Operation(length, {f"{qubit.name}.Gamma_down": CTerm(qubit.a, coeffs=coeff)})

    def t1_sequence(max_time=3000, pulsed_t1=None):
        seq = get_sequence(system)
        if pulsed_t1 is not None:
            lossy_pi_pulse(qubit, pulsed_t1)
            sync()
            lossy_delay(qubit, max_time, pulsed_t1)
        else:
            qubit.rotate_x(np.pi)
            sync()
            delay(max_time)
        return seq

    def fit_tau(ys, t0):
        ts = np.arange(len(ys))
        slope, _ = np.polyfit(ts[t0:], np.log(ys[t0:]), 1)
        return -1 / slope

    t0 = qubit.gaussian_pulse.sigma * qubit.gaussian_pulse.chop
    res = t1_sequence().run(qubit.fock(0), e_ops=[qubit.fock(1)], only_final_state=False)
    assert len(res.states) == t0 + 3000
    assert abs(fit_tau(res.expect[0], t0) - qubit.t1) < 0.1
    for factor in (2, 5):
        pulsed = qubit.t1 / factor
        res = t1_sequence(pulsed_t1=pulsed).run(qubit.fock(0), e_ops=[qubit.fock(1)], only_final_state=False)
        assert abs(fit_tau(res.expect[0], t0) - pulsed) < 0.1


def test_sequence_with_unitaries(kat_engine):
    """test_sequencing.py:343-408: interleaved pulses and instantaneous unitaries, N=30."""
    qubit = Transmon("qubit", levels=3, kerr=-200e-3)
    cavity = Cavity("cavity", levels=10, kerr=-10e-6)
    system = System("system", modes=[qubit, cavity])
    system.set_cross_kerr(cavity, qubit, chi=-2e-3)
    init = system.ground_state()
    n_rot = 6
    theta = np.pi / n_rot
    seq = Sequence(system)
    for _ in range(n_rot):
        qubit.rotate_x(theta / 2)
        seq.capture()
        seq.append(qubit.Rx(theta / 2))
    result = seq.run(init, e_ops=[qubit.fock_dm(1)])
    assert len(result.states) == result.times.size
    assert qt.fidelity(result.states[-1], qubit.Rx(np.pi) * init) ** 2 > 0.999
    assert result.expect[0][-1] > 0.99

    seq = Sequence(system)
    for _ in range(n_rot):
        seq.append(qubit.rotate_x(theta / 4, capture=False))
        sync()
        seq.append(qubit.rotate_x(theta / 4, capture=False))
        seq.append(qubit.Rx(theta / 4))
        seq.append(qubit.Rx(theta / 4))
    props = seq.propagator()
    assert qt.fidelity(props[-1] * init, qubit.Rx(np.pi) * init) ** 2 > 0.9995


def test_rotations_grid(kat_engine):
    """test_rotations.py: pulsed rotate(theta, phi) vs Raxis, F > 0.999 (3x3 of the 11x11 grid)."""
    qubit = Transmon("qubit", levels=2)
    system = System("system", modes=[qubit])
    tune_rabi(system, system.ground_state(), plot=False, verify=False, progbar=False)
    init = qubit.fock(0)
    for theta in np.linspace(-np.pi, np.pi, 3):
        for phi in np.linspace(0, 2 * np.pi, 3):
            seq = Sequence(system)
            qubit.rotate(theta, phi)
            result = seq.run(init)
            assert qt.fidelity(result.states[-1], qubit.Raxis(theta, phi) * init) ** 2 > 0.999
