"""Ad-hoc GPU check: DMMA / generic kernels vs the oracle on the named workloads + timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import oracle_solve, random_problem, trace_distance
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib

eng = Engine.get(0)
which = sys.argv[1:] or ["rand", "c1", "c3", "c2", "c4"]

def cmp(name, bp, methods, n_check=2, oracle=True):
    outs = {}
    for m in methods:
        bp.method = m
        t = time.time(); out = eng.solve(bp); wall = time.time() - t
        outs[m] = out
        print(f"[{name}] method={_lib.METHOD_NAMES[m]:8s} N={bp.N} B={bp.B} status_ok={bool(np.all(out.status==0))} kernel_ms={out.kernel_ms:.2f} wall={wall:.2f}s "
              f"steps(acc/rej)={out.stats[:,0].mean():.1f}/{out.stats[:,1].mean():.1f} launches={out.launches}", flush=True)
    ms = list(outs)
    for a in ms[1:]:
        d = np.max(np.abs(outs[ms[0]].final_state - outs[a].final_state)); de = np.max(np.abs(outs[ms[0]].expect - outs[a].expect)) if outs[a].expect.size else 0
        print(f"[{name}]   {_lib.METHOD_NAMES[ms[0]]} vs {_lib.METHOD_NAMES[a]}: max|dfinal|={d:.2e} max|dexpect|={de:.2e}")
    if oracle:
        for b in range(min(n_check, bp.B)):
            t = time.time(); ref = oracle_solve(bp, b); to = time.time() - t
            for m in ms:
                td = trace_distance(outs[m].final_state[b], ref.final_state)
                ee = max([np.max(np.abs(outs[m].expect[b, i] - np.real(ref.expect[i]))) for i in range(len(ref.expect))] or [0])
                print(f"[{name}]   b={b} {_lib.METHOD_NAMES[m]:8s} vs oracle: td={td:.2e} expect={ee:.2e}  (oracle {to:.2f}s, nfev={ref.nfev})", flush=True)

G, D, A = _lib.METHOD_GENERIC, _lib.METHOD_DMMA, _lib.METHOD_AUTO
if "rand" in which:
    for N in (3, 8, 12, 20, 33, 45):
        cmp(f"rand{N}", random_problem(N=N, B=4, n_t=20, n_ch=2, n_c=2, n_ctd=1, n_e=2, seed=N, hscale=0.3), [G, D], oracle=(N <= 20))
if "c1" in which:
    cmp("c1", w.c1_rabi(51), [G, D])
if "c3" in which:
    cmp("c3", w.c3_drag(16), [G, D])
if "c2" in which:
    cmp("c2", w.c2_selective(4, 4), [G, D])
S = _lib.METHOD_DMMA_STREAM
if "stream" in which:
    for N in (49, 57, 64, 65, 72, 80, 81, 90, 96):
        cmp(f"rand{N}", random_problem(N=N, B=2, n_t=6, n_ch=2, n_c=2, n_ctd=1, n_e=2, seed=N, hscale=0.3), [G, S], oracle=False)
if "c4" in which:
    cmp("c4", w.c4_two_transmons_bus(4), [G, S], n_check=1)
if "c4big" in which:
    cmp("c4big", w.c4_two_transmons_bus(148), [S], oracle=False)
if "c2big" in which:
    bp = w.c2_selective(16, 16)
    cmp("c2big", bp, [D], oracle=False)
