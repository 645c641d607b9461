"""Ad-hoc GPU check of SEQ_METHOD_SPARSE: row-sparse random problems and the named workloads against
the FP64-FMA generic kernel (states, expectation traces, step counts), plus C2 timing vs the DMMA kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import row_sparse_problem
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib

eng = Engine.get(0)
which = sys.argv[1:] or ["rand", "auto", "c2s", "c4", "time"]
S, G, D, A, DG = _lib.METHOD_SPARSE, _lib.METHOD_GENERIC, _lib.METHOD_DMMA, _lib.METHOD_AUTO, _lib.METHOD_DIAG
bad = 0

def cmp(name, mk, ref=G, tol=1e-12, methods=(S, DG)):
    global bad
    g = mk(); g.method = ref
    og = eng.solve(g)
    for m in methods:
        a = mk(); a.method = m
        try:
            oa = eng.solve(a)
        except RuntimeError as ex:
            print(f"[{name}] {_lib.METHOD_NAMES[m]}: {ex}"); continue
        ds = np.max(np.abs(oa.states - og.states)) if oa.states is not None else np.max(np.abs(oa.final_state - og.final_state))
        df = np.max(np.abs(oa.final_state - og.final_state))
        de = np.max(np.abs(oa.expect - og.expect)) if oa.expect.size else 0.0
        same = np.array_equal(oa.stats[:, :2], og.stats[:, :2])
        ok = oa.method == _lib.METHOD_NAMES[m] and np.all(oa.status == 0) and ds < tol and df < tol and de < tol and same
        bad += 0 if ok else 1
        print(f"[{name}] {'OK ' if ok else 'BAD'} method={oa.method} vs {og.method}: states={ds:.1e} final={df:.1e} expect={de:.1e} stats_equal={same} "
              f"steps={oa.stats[:, 0].mean():.0f}/{oa.stats[:, 1].mean():.0f} kernel_ms={oa.kernel_ms:.2f} vs {og.kernel_ms:.2f}", flush=True)

if "rand" in which:
    for N, w_h, w_l, n_c, n_ctd in ((7, 3, 1, 2, 0), (12, 3, 1, 2, 1), (20, 5, 2, 2, 1), (31, 3, 1, 3, 0), (45, 3, 1, 4, 0), (46, 5, 1, 2, 0),
                                    (47, 3, 2, 2, 0), (48, 3, 1, 1, 1), (49, 3, 1, 2, 0), (63, 5, 3, 1, 0), (64, 3, 1, 2, 0), (79, 3, 1, 2, 1),
                                    (80, 5, 1, 2, 0), (90, 5, 1, 3, 0), (100, 3, 1, 2, 0), (112, 3, 2, 1, 0), (45, 3, 1, 0, 0)):
        cmp(f"rand N={N} wh={w_h} wl={w_l} nc={n_c}+{n_ctd}",
            lambda: row_sparse_problem(N=N, B=3, n_t=8, w_h=w_h, w_l=w_l, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N, hscale=0.3, store_states=True))
    # uncapped steps (rejections), complex expectation, shared table + coeff_scale
    def mk():
        bp = row_sparse_problem(N=24, B=4, n_t=12, w_h=3, w_l=1, n_ch=2, n_c=2, n_e=2, seed=5, hscale=3.0, store_states=True)
        rng = np.random.default_rng(2)
        bp.E = bp.E + 0.3 * (rng.normal(size=bp.E.shape) + 1j * rng.normal(size=bp.E.shape))
        bp.expect_complex = True
        bp.coeff = bp.coeff[:1]
        bp.coeff_scale = rng.uniform(0.2, 3.0, size=(4, 2))
        bp.max_step = 0.0
        return bp
    cmp("uncapped/complex/scale", mk, tol=1e-11)
if "auto" in which:
    for nm, bp in (("c2", w.c2_selective(2, 2)), ("c4", w.c4_two_transmons_bus(2)), ("c2small", w.c2_selective(1, 2, sigma=10, cavity_levels=4))):
        print("[auto]", nm, eng.solve(bp).method)
if "c2s" in which:
    cmp("c2 small", lambda: w.c2_selective(2, 2, sigma=10, cavity_levels=4))
    cmp("c2 full", lambda: w.c2_selective(2, 2), ref=D, tol=1e-12)
if "c4" in which:
    cmp("c4", lambda: w.c4_two_transmons_bus(2), ref=_lib.METHOD_DMMA_STREAM)
if "time" in which:
    for B in (148, 296, 592):
        bp = w.c2_selective(B // 4, 4)
        for m in (DG, S, D):
            bp.method = m
            eng.solve(bp)
            out = eng.solve(bp)
            print(f"[time c2] B={bp.B} method={out.method}: kernel {out.kernel_ms:.1f} ms -> {bp.B / out.kernel_ms * 1e3:.0f} traj/s (steps {out.stats[:, 0].mean():.0f}, rhs {out.stats[:, 2].mean():.0f})", flush=True)
    bp = w.c4_two_transmons_bus(148)
    for m in (DG, S, _lib.METHOD_DMMA_STREAM):
        bp.method = m
        out = eng.solve(bp)
        print(f"[time c4] B={bp.B} method={out.method}: kernel {out.kernel_ms:.1f} ms -> {bp.B / out.kernel_ms * 1e3:.0f} traj/s (steps {out.stats[:, 0].mean():.0f}, rhs {out.stats[:, 2].mean():.0f})", flush=True)
print("SPARSE_CHECK", "FAILED" if bad else "OK", bad)
if "fast" in which:
    # C2-shaped systems of several sizes (FAST variant of the diagonal kernel: rank-1 drive diagonals) vs the generic kernel
    for lv in (4, 7, 10, 15, 16):
        mk = lambda: w.c2_selective(2, 2, sigma=10, cavity_levels=lv)
        bp = mk()
        out = eng.solve(bp)
        print("[fast] levels", lv, "N", bp.N, "->", out.method, "|", out.variant)
        cmp(f"c2 levels={lv}", mk, methods=(DG,))
