"""Global error of the engine at the reference's default tolerances as a function of the order-switch thresholds
(SolveParams::stay_err / probe_err; SEQ_STAY_ERR / SEQ_PROBE_ERR in the environment), against the engine's own converged
run (atol 1e-13, rtol 1e-11, no order switching) on the named workloads, with the kernel time next to it.
    python tools/acc_sweep.py [c1 c2 c3 c4 ket]"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib

eng = Engine.get(0)


def tdist(a, b):
    d = a - b
    if d.ndim == 3:
        d = 0.5 * (d + np.conj(np.swapaxes(d, 1, 2)))
        return 0.5 * np.abs(np.linalg.eigvalsh(d)).sum(axis=1)
    # kets: distance of the projectors = sqrt(1 - |<a|b>|^2)
    ov = np.abs(np.sum(np.conj(a) * b, axis=1)) ** 2
    return np.sqrt(np.maximum(0.0, 1.0 - ov))


def ket243(B=64):
    from test_rows_n1_kets import _five_qutrits, _lowered
    from sequencing_b200 import qobj as qt
    from sequencing_b200.sequencing import sync
    system, qs = _five_qutrits()
    rng = np.random.default_rng(6)
    psi = np.zeros(243, complex); psi[0] = 1.0
    psi[rng.integers(0, 243, 6)] += 0.3 * (rng.normal(size=6) + 1j * rng.normal(size=6))
    psi /= np.linalg.norm(psi)
    init = qt.Qobj(psi.reshape(-1, 1), dims=[[3] * 5, [1] * 5])
    def build():
        qs[0].rotate_y(np.pi / 2); qs[3].rotate_x(np.pi / 3); sync(); qs[1].rotate_y(np.pi)
    bp = _lowered(system, build, init)
    bp.E = np.stack([qs[0].n.full(), qs[1].n.full()]).astype(complex)
    return bp


makers = {"c1": lambda: w.c1_rabi(4096), "c2": lambda: w.c2_selective(32, 32), "c3": lambda: w.c3_drag(512), "c4": lambda: w.c4_two_transmons_bus(512), "ket": ket243}
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "ket"]
for name in which:
    bp = makers[name]()
    for k in ("SEQ_STAY_ERR", "SEQ_PROBE_ERR"):
        os.environ.pop(k, None)
    t = copy.copy(bp)
    t.atol, t.rtol, t.nsteps = 1e-13, 1e-11, 1000000
    t.order_switch = False
    tru = eng.solve(t)
    assert np.all(tru.status == 0), tru.status
    for stay, probe in ((0.5, 0.02), (0.1, 0.02), (0.05, 0.01), (0.02, 0.005), (0.01, 0.002), (0.003, 0.001)):
        os.environ["SEQ_STAY_ERR"], os.environ["SEQ_PROBE_ERR"] = str(stay), str(probe)
        eng.solve(bp)
        out = eng.solve(bp)
        td = tdist(out.final_state, tru.final_state)
        ee = np.abs(out.expect - tru.expect).reshape(bp.B, -1).max(axis=1)
        print(f"{name:4s} stay {stay:<6} probe {probe:<6} kernel_ms {out.kernel_ms:9.3f}  rhs mean {out.stats[:, 2].mean():8.1f}  td max {td.max():.2e} p95 {np.percentile(td, 95):.2e}  "
              f"expect max {ee.max():.2e} p95 {np.percentile(ee, 95):.2e}  frac<=1e-6: td {np.mean(td <= 1e-6):.3f} exp {np.mean(ee <= 1e-6):.3f}", flush=True)
