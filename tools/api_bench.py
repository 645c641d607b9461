"""End-to-end numbers through the PUBLIC API (what a user of the reference calls), wall clock, host lowering included:
  * calibration.tune_rabi  (C1 shape: transmon(3), T1/T2)   4 096 and 65 536 sweep points   (calibration.py:65-169)
  * calibration.tune_drag  (C3 shape: transmon(4))          4 096 DRAG values x 2 sequences  (calibration.py:396-503)
  * Sequence / run_sequences (C4 as SURVEY 8(d) defines it): 512 sequences [pulsed R_y(pi/2) on both qubits, CX-type
    unitary, pulsed R_y with amplitude / phase jitter] on two transmons(3) + bus cavity(10), full collapse set  (main.py:250-340)
  * kets: 1 024 five-qutrit ket trajectories (N = 243, the sesolve path; test_qasm.py:463-530 shape) through the engine
One JSON line per item.    python tools/api_bench.py [rabi drag c4seq ket]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sequencing_b200 import Cavity, System, Transmon, get_sequence, sync
from sequencing_b200 import qobj as qt
from sequencing_b200.calibration import tune_drag, tune_rabi
from sequencing_b200.sequencing.containers import Sequence, run_sequences
from sequencing_b200.solver import Engine
from sequencing_b200.system import CouplingTerm

which = sys.argv[1:] or ["rabi", "drag", "c4seq", "ket"]
eng = Engine.get(0)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t)
    return min(ts), out


if "rabi" in which:
    for n in (4096, 65536):
        for dev in (None, False) if n == 4096 else (None,):
            def go():
                q = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
                system = System("system", modes=[q])
                return tune_rabi(system, q.fock(0), amp_range=(-2, 2, n), progbar=False, plot=False, update=False, verify=False, device_synthesis=dev)
            dt, (_, old, new) = timed(go)
            print(json.dumps({"item": "tune_rabi", "points": n, "device_synthesis": "default (on)" if dev is None else "off (per-point host compilation)",
                              "wall_s": dt, "sequences_per_s": n / dt, "new_amp": new}), flush=True)

if "drag" in which:
    n = 4096
    def go():
        q = Transmon("qubit", levels=4, kerr=-0.2, t1=20e3, t2=15e3)
        system = System("system", modes=[q])
        return tune_drag(system, q.fock(0), drag_range=(-5, 5, n), progbar=False, plot=False, update=False)
    dt, (_, old, new) = timed(go)
    print(json.dumps({"item": "tune_drag", "drag_values": n, "sequences": 2 * n, "wall_s": dt, "sequences_per_s": 2 * n / dt, "new_drag": new}), flush=True)

if "c4seq" in which:
    B = 512
    common = dict(levels=3, kerr=-0.2, t1=20e3, t2=15e3, thermal_population=0.01)
    q0, q1 = Transmon("qubit0", **common), Transmon("qubit1", **common)
    cav = Cavity("cavity", levels=10, kerr=-1e-5, t1=100e3, t2=80e3, thermal_population=0.01)
    system = System("system", modes=[q0, q1, cav])
    system.set_cross_kerr(cav, q0, chi=-2e-3)
    system.set_cross_kerr(cav, q1, chi=-1.5e-3)
    system.set_cross_kerr(q0, q1, chi=-1e-4)
    system.coupling_terms[frozenset([q0.name, cav.name])].append(CouplingTerm([(q0, "a"), (cav, "ad")], strength=2 * np.pi * 1e-3, add_hc=True))
    # CX on the qubit subspaces of (q0, q1), identity elsewhere (the closed-form gate a qasm `CX` line stands for)
    cx = np.eye(9, dtype=complex)
    i10, i11 = 1 * 3 + 0, 1 * 3 + 1
    cx[[i10, i11]] = cx[[i11, i10]]
    CX = qt.Qobj(np.kron(cx, np.eye(10)), dims=[[3, 3, 10], [3, 3, 10]])
    rng = np.random.default_rng(0)
    jit = rng.normal(0, 1e-2, size=(B, 2))
    init = system.fock()

    def go():
        seqs = []
        for da, dp in jit:
            sq = Sequence(system)
            with q0.pulse_scale(1 + da), q1.pulse_scale(1 + da):
                q0.rotate(np.pi / 2, np.pi / 2 + dp)
                q1.rotate(np.pi / 2, np.pi / 2 - dp)
            sq.capture()
            sq.append(CX)
            with q1.pulse_scale(1 + da):
                q1.rotate(np.pi / 3, np.pi / 2 + dp)
            sq.capture()
            seqs.append(sq)
        return run_sequences(seqs, [init] * B, e_ops=[q0.fock_dm(1), q1.fock_dm(1)])
    dt, res = timed(go, reps=2)
    print(json.dumps({"item": "C4 as a Sequence: pulsed segment, CX unitary, pulsed segment (run_sequences)", "B": B, "N": 90, "wall_s": dt,
                      "sequences_per_s": B / dt, "expect_q1_final_mean": float(np.mean([r.expect[1][-1] for r in res]))}), flush=True)

if "ket" in which:
    from test_rows_n1_kets import _five_qutrits, _lowered
    system, qs = _five_qutrits()
    rng = np.random.default_rng(6)
    psi = np.zeros(243, complex); psi[0] = 1.0
    psi[rng.integers(0, 243, 6)] += 0.3 * (rng.normal(size=6) + 1j * rng.normal(size=6))
    psi /= np.linalg.norm(psi)
    init = qt.Qobj(psi.reshape(-1, 1), dims=[[3] * 5, [1] * 5])
    def build():
        qs[0].rotate_y(np.pi / 2); qs[3].rotate_x(np.pi / 3); sync(); qs[1].rotate_y(np.pi)
    bp = _lowered(system, build, init)
    B = 1024
    bp.state0 = np.ascontiguousarray(np.repeat(bp.state0, B, axis=0))
    bp.coeff = np.ascontiguousarray(np.repeat(bp.coeff, B, axis=0) * (1 + 1e-3 * rng.normal(size=(B, 1, 1))))
    bp.E = np.stack([qs[0].n.full(), qs[1].n.full()]).astype(complex)
    eng.solve(bp)
    out = eng.solve(bp)
    print(json.dumps({"item": "ket path (sesolve semantics), five qutrits", "N": 243, "B": B, "n_t": int(bp.coeff.shape[-1]), "method": out.method,
                      "kernel_ms": out.kernel_ms, "trajectories_per_s": B / (out.kernel_ms * 1e-3), "rhs_mean": float(out.stats[:, 2].mean())}), flush=True)
