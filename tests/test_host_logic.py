"""Host-side logic that needs no GPU: channel compiler semantics, containers, Qobj stand-in,
parameter trees, lowering/batching, and the C-ABI library (loads, exports, struct layout)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import sequencing_b200.qobj as qt
from sequencing_b200 import (Cavity, System, Transmon, delay, delay_channels, get_sequence, ket2dm, ops2dms, sync)
from sequencing_b200.pulses import GaussianPulse, gaussian_chop_norm, gaussian_wave
from sequencing_b200.sequencing import HamiltonianChannels, ValidatedList
from sequencing_b200.solver import Options, _lib, deferred_solves, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- reference test_sequencing.py:28-119 ------------------------------------------------
def test_ket2dm_ops2dms():
    assert ket2dm(qt.fock(5, 0)) == qt.fock_dm(5, 0)
    assert ket2dm(qt.fock_dm(3, 0)) == qt.fock_dm(3, 0)
    for n, dm in enumerate(ops2dms([qt.fock(5, n) for n in range(5)])):
        assert dm == qt.fock_dm(5, n)


def test_validated_list():
    class Numbers(ValidatedList):
        VALID_TYPES = (int, float, complex)

    inst = Numbers(iterable=[1, 1j, 0.5])
    with pytest.raises(TypeError):
        inst.append("a string")
    with pytest.raises(TypeError):
        inst._validate([1, 2, 3])
    inst.extend([0, float("inf"), -1j])
    assert inst.pop() == -1j


def test_hamiltonian_channels_errors_and_delays():
    hc = HamiltonianChannels()
    with pytest.raises(ValueError):
        hc.add_channel("H0")
    hc.add_channel("H0", H=qt.qeye(3), time_dependent=False)
    with pytest.raises(ValueError):
        hc.add_channel("H0", H=qt.qeye(3), time_dependent=False, error_if_exists=True)
    with pytest.raises(ValueError):
        hc.add_operation("H0", t0=0, duration=100, H=None, C_op=None)
    with pytest.raises(ValueError):
        hc.add_operation("H0", t0=0, duration=100, H=qt.qeye(3))   # H0 is time-independent
    with pytest.raises(ValueError):
        hc.add_operation("H1", H=qt.qeye(3))                        # neither times nor t0/duration
    with pytest.raises(ValueError):
        hc.delay_channels("H3", 5)
    with pytest.raises(TypeError):
        hc.add_channel("bad", H=np.eye(3))
    hc.add_operation("H2", H=qt.qeye(3), t0=0, duration=10)
    hc.delay_channels("H1", 5)
    hc.add_operation("H1", t0=0, duration=10)
    assert hc.channels["H1"]["delay"] == 5 and hc.channels["H2"]["delay"] == 0
    H, C_ops, times = hc.build_hamiltonian()
    assert len(H) == 3 and C_ops == [] and times.size == 15
    # H1 is shifted by its delay, H2 is right-padded with zeros
    np.testing.assert_array_equal(hc.channels["H1"]["coeffs"], np.r_[np.zeros(5), np.ones(10)])
    np.testing.assert_array_equal(hc.channels["H2"]["coeffs"], np.r_[np.ones(10), np.zeros(5)])


def test_same_channel_operations_sum_and_callable_coeffs():
    hc = HamiltonianChannels()
    hc.add_operation("a", H=qt.qeye(2), t0=0, duration=4, coeffs=2.0)
    hc.add_operation("a", t0=2, duration=4, coeffs=lambda t, k=1.0: k * t, coeffs_kwargs={"k": 0.5, "reset_t0": True})
    np.testing.assert_allclose(hc.channels["a"]["coeffs"], [2, 2, 2 + 0, 2 + 0.5, 1.0, 1.5])
    assert hc.times.tolist() == [0, 1, 2, 3, 4, 5]
    # an all-zero collapse table is emitted as a constant operator (basic.py:242-247)
    hc.add_operation("c", C_op=qt.destroy(2), t0=0, duration=6, coeffs=0.0)
    _, C_ops, _ = hc.build_hamiltonian()
    assert isinstance(C_ops[0], qt.Qobj)


@pytest.fixture
def qc_system():
    qubit = Transmon("qubit", levels=3, kerr=-200e-3)
    cavity = Cavity("cavity", levels=10, kerr=-10e-6)
    system = System("system", modes=[qubit, cavity])
    system.set_cross_kerr(cavity, qubit, chi=-2e-3)
    return system


# ---- reference test_sequencing.py:260-331 -----------------------------------------------
def test_sync_delay_and_delay_channels_timing(qc_system):
    system = qc_system
    qubit, cavity = system.qubit, system.cavity
    n = qubit.gaussian_pulse.sigma * qubit.gaussian_pulse.chop
    seq = get_sequence(system)
    qubit.rotate_x(np.pi / 2)
    qubit.rotate_x(np.pi / 2)
    assert all(ch["coeffs"].size == n for ch in seq.channels.values())
    seq = get_sequence(system)
    qubit.rotate_x(np.pi / 2); sync(); qubit.rotate_x(np.pi / 2)
    assert all(ch["coeffs"].size == 2 * n for ch in seq.channels.values())
    seq = get_sequence(system)
    qubit.rotate_x(np.pi / 2); delay(100); qubit.rotate_x(np.pi / 2)
    assert all(ch["coeffs"].size == 100 + 2 * n for ch in seq.channels.values())
    seq = get_sequence(system)
    qubit.rotate_x(np.pi)
    delay_channels({"cavity.x": cavity.x, "cavity.y": cavity.y}, 5)
    cavity.displace(1)
    assert all(ch["coeffs"].size == 5 + n for ch in seq.channels.values())
    seq = get_sequence(system)
    cavity.displace(1)
    delay_channels({"qubit.x": qubit.x, "qubit.y": qubit.y}, 5)
    qubit.rotate_x(np.pi)
    assert all(ch["coeffs"].size == 5 + n for ch in seq.channels.values())


def test_system_operators_and_ordering(qc_system):
    system = qc_system
    assert [m.name for m in system.modes] == ["qubit", "cavity"]
    assert system.I().shape == (30, 30) and system.I().dims == [[3, 10], [3, 10]]
    assert len(system.H0()) == 3 and len(system.c_ops()) == 0      # qubit kerr, cavity kerr, cross-kerr
    system.qubit.t1, system.qubit.t2 = 20e3, 15e3
    assert len(system.c_ops()) == 2                                  # decay + dephasing, zero operators dropped
    chi_term = system.couplings()[0]
    assert chi_term == 2 * np.pi * (-2e-3) * system.qubit.n * system.cavity.n
    a = system.qubit.a
    assert a * a.dag() - a.dag() * a != 0
    with system.use_modes([system.qubit]):
        assert system.qubit.a.shape == (3, 3)
    assert system.qubit.a.shape == (30, 30)
    roundtrip = System.from_json(json_str=system.to_json(dumps=True))
    assert roundtrip.cross_kerrs == system.cross_kerrs and len(roundtrip.H0()) == 3


def test_waveform_constants_and_pulse_call_precedence():
    # gaussian_chop_norm(10, 4) = 42.819716308728
    assert abs(gaussian_chop_norm(10, 4) - 42.819716308728) < 1e-9
    w = gaussian_wave(10, chop=4)
    assert w.size == 40 and w[0] == 0 and w[-1] == 0 and abs(w.max() - 1) < 2e-3
    p = GaussianPulse(name="g")
    c = p(amp=0.5, phase=np.pi / 2)
    assert c.dtype == complex and c.size == 40
    np.testing.assert_allclose(c.real, 0, atol=1e-15)
    p.dt = 0.5
    assert p().size == 80                       # dt feeds the shape function
    q = Transmon("qubit", levels=2)
    op = q.rotate_x(np.pi, capture=False)
    assert op.duration == 40 and set(op.terms) == {"qubit.x", "qubit.y"}
    cav = Cavity("cavity", levels=5)
    opc = cav.displace(1j, capture=False)
    np.testing.assert_allclose(opc.terms["cavity.x"].coeffs, -cav.gaussian_pulse(amp=2j / gaussian_chop_norm(10, 4)).imag)


def test_qobj_subset():
    a = qt.destroy(4)
    assert (a.dag() * a) == qt.num(4)
    assert qt.tensor(qt.qeye(2), a).dims == [[2, 4], [2, 4]]
    psi = (qt.basis(2, 0) + 1j * qt.basis(2, 1)).unit()
    assert abs(psi.norm() - 1) < 1e-15 and psi.isket and not psi.isoper
    rho = qt.ket2dm(psi)
    assert rho.isherm and abs(rho.tr() - 1) < 1e-15
    assert abs(qt.fidelity(rho, psi) - 1) < 1e-12 and qt.tracedist(rho, rho) < 1e-15
    assert abs(qt.expect(qt.sigmaz(), psi)) < 1e-15
    U = (-1j * np.pi / 2 * qt.sigmax()).expm()
    assert abs(qt.fidelity(U * qt.basis(2, 0), qt.basis(2, 1)) - 1) < 1e-12
    big = qt.tensor(qt.fock_dm(2, 1), qt.fock_dm(3, 2))
    assert big.ptrace(1) == qt.fock_dm(3, 2) and big.ptrace(0) == qt.fock_dm(2, 1)
    assert qt.displace(20, 0.5).dag() * qt.displace(20, 0.5) == qt.qeye(20)
    assert (0 * a).data.nnz == 0 and a.data.nnz == 3


# ---- lowering and batching ------------------------------------------------------------------
def test_lowering_shapes_and_semantics():
    qubit = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
    system = System("system", modes=[qubit])
    seq = get_sequence(system)
    qubit.rotate_x(np.pi)
    cs = seq.compile()
    H, times = cs._assemble([])
    opt = Options(); opt.max_step = 1
    it = api._lower(H, qubit.fock(0), times, system.c_ops(), [qubit.fock_dm(1)], opt)
    assert not it.is_ket and it.state0.shape == (3, 3)          # ket + collapse ops -> density matrix
    assert len(it.Hk) == 2 and len(it.Lc) == 2 and it.n_t == 40 and it.out_idx.tolist() == list(range(40))
    np.testing.assert_allclose(it.H0, np.pi * (-0.2) * np.diag([0, 0, 2.0]))
    it2 = api._lower(H, qubit.fock(0), times, [], [], opt)
    assert it2.is_ket and it2.store_states                       # no e_ops -> states are stored
    evo = api.QobjEvo(H, tlist=times)
    it3 = api._lower(evo, qubit.fock_dm(0), [times.min(), times.max()], [], [], opt)
    assert it3.out_idx.tolist() == [0, 39] and it3.n_t == 40
    # complex coefficient -> two real channels; |c|^2 for collapse terms
    c = np.exp(1j * np.linspace(0, 1, 40))
    it4 = api._lower([[qubit.x, c]], qubit.fock_dm(0), times, [[qubit.a, 0.1 * c]], [], opt)
    assert len(it4.Hk) == 2 and np.allclose(it4.Hk[1], 1j * it4.Hk[0])
    np.testing.assert_allclose(it4.rate[0], 0.01)
    with pytest.raises(ValueError):
        api._lower([[qubit.x, c[:5]]], qubit.fock_dm(0), times, [], [], opt)
    with pytest.raises(ValueError):
        api._lower([[qubit.x, c]], qubit.fock_dm(0), [0.0, 0.5], [], [], opt)


def test_deferred_solves_group_into_batches(oracle_engine):
    qubit = Transmon("qubit", levels=2)
    system = System("system", modes=[qubit])
    results = []
    with deferred_solves():
        for amp in (0.5, 1.0, 1.5):
            seq = get_sequence(system)
            with qubit.pulse_scale(amp):
                qubit.rotate_x(np.pi)
            results.append(seq.run(qubit.fock_dm(0), e_ops=[qubit.fock_dm(1)]))
        seq = get_sequence(system)                       # different length -> its own batch
        qubit.rotate_x(np.pi); sync(); qubit.rotate_x(np.pi)
        results.append(seq.run(qubit.fock_dm(0), e_ops=[qubit.fock_dm(1)]))
        assert results[0]._pending
    assert sorted(oracle_engine.batch_sizes) == [1, 3]
    assert not results[0]._pending and results[1].expect[0][-1] > 0.99 and len(results[3].states) == 80
    assert results[0].num_expect == 1 and results[0].solver == "mesolve"


def test_frozen_system_evaluation_cache_and_signature_digests(qc_system):
    """``System.frozen()`` (what ``run_sequences`` lowers a batch under): equal system STATES - also deep copies, which
    every captured segment carries (``main.py:226-233``) - share one evaluation of H0 / c_ops as the same objects; a
    different state gets its own; outside the block nothing is cached.  The deferred-solve signature hashes each source
    operator once per object and still groups by CONTENT (equal operators of different identity share a batch, a changed
    operator does not)."""
    import copy
    system = qc_system
    twin = copy.deepcopy(system)
    ref_h0 = [h.full() for h in system.H0()]
    with system.frozen():
        a, b = system.H0(), twin.H0()
        assert all(x is y for x, y in zip(a, b)) and len(a) == len(b) == len(ref_h0)
        assert all(x is y for x, y in zip(system.c_ops(), twin.c_ops()))
        assert system.H0_sum() is twin.H0_sum()
        other = copy.deepcopy(system)
        other.modes[0].kerr = other.modes[0].kerr * 1.5
        assert other.H0_sum() is not system.H0_sum()
        assert not np.allclose(other.H0_sum().full(), system.H0_sum().full())
    assert all(np.array_equal(x.full(), y) for x, y in zip(a, ref_h0))
    assert system.H0_sum() is not system.H0_sum()      # (no cache outside the block)
    # signatures: same content, different identity -> one group; different content -> two
    psi = system.fock()
    with deferred_solves(execute=False) as d:
        for sysk in (system, twin, other):
            H = [sum(sysk.H0())]
            api.mesolve(H, psi, np.arange(4.0), sysk.c_ops(), [])
    groups = sorted(len(g) for g in d.groups())
    assert groups == [1, 2]


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from sequencing_b200.solver import EngineUnavailable
    qubit = Transmon("qubit", levels=2)
    system = System("system", modes=[qubit])
    seq = get_sequence(system)
    qubit.rotate_x(np.pi)
    with pytest.raises(EngineUnavailable):
        seq.run(qubit.fock(0))


# ---- the C-ABI library ----------------------------------------------------------------------
def test_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "seq_engine.h")).read()
    declared = set(re.findall(r"\b(seq_engine_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.seq_engine_abi_version() == _lib.ABI_VERSION == 3
    assert b"sm_100a" in lib.seq_engine_build_info()


def test_abi_struct_layout_matches_header(tmp_path):
    src = tmp_path / "layout.c"
    fields = [f for f, _ in _lib.SolveDesc._fields_]
    body = "".join(f'printf("{f} %zu\\n", offsetof(seq_solve_desc, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "seq_engine.h"\nint main(){printf("__sizeof %zu\\n", sizeof(seq_solve_desc));' + body + "return 0;}")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["__sizeof"]) == ctypes.sizeof(_lib.SolveDesc)
    for f in fields:
        assert int(out[f]) == getattr(_lib.SolveDesc, f).offset, f


def test_abi_validation_without_gpu():
    """Descriptor validation and method selection are host-only: callable without a device."""
    lib = _lib.load()
    d = _lib.SolveDesc()
    d.size = ctypes.sizeof(_lib.SolveDesc)
    d.N, d.B, d.n_t, d.n_out, d.n_c = 45, 8, 100, 100, 4
    assert lib.seq_engine_select_method(None, ctypes.byref(d)) == _lib.METHOD_DMMA
    d.flags = _lib.FLAG_KET
    assert lib.seq_engine_select_method(None, ctypes.byref(d)) == _lib.METHOD_GENERIC
    d.flags = 0
    for N, want in ((3, _lib.METHOD_SMALL), (6, _lib.METHOD_SMALL), (7, _lib.METHOD_DMMA), (48, _lib.METHOD_DMMA),
                    (49, _lib.METHOD_DMMA_STREAM), (96, _lib.METHOD_DMMA_STREAM), (97, _lib.METHOD_LARGE), (1875, _lib.METHOD_LARGE)):
        d.N = N
        assert lib.seq_engine_select_method(None, ctypes.byref(d)) == want, N
    d.N = 45
    assert lib.seq_engine_workspace_bytes(None, ctypes.byref(d)) > 0
    d.size = 7
    assert lib.seq_engine_workspace_bytes(None, ctypes.byref(d)) == -1
