"""Error attribution on long sequences: GPU (default/tight) vs oracle (default/tight/very tight)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, copy
from helpers import oracle_solve, trace_distance
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib
eng = Engine.get(0)
def run(name, bp, nb=2):
    outs = {}
    for tag, (a, r) in {"def": (1e-8, 1e-6), "tight": (1e-10, 1e-8), "vtight": (1e-12, 1e-10)}.items():
        q = copy.copy(bp); q.atol, q.rtol = a, r
        outs[tag] = eng.solve(q)
        print(name, tag, "gpu steps/rhs", outs[tag].stats[:nb, :3].tolist(), "ms", outs[tag].kernel_ms)
    for b in range(nb):
        refs = {}
        for tag, (a, r) in {"def": (1e-8, 1e-6), "tight": (1e-10, 1e-8), "vtight": (1e-13, 1e-11)}.items():
            t = time.time(); refs[tag] = oracle_solve(bp, b, atol=a, rtol=r, nsteps=100000); 
            print(f"  oracle {tag}: {time.time()-t:.2f}s nfev={refs[tag].nfev}")
        truth = refs["vtight"]
        def d(x, y): return trace_distance(x, y)
        def de(o, r): return max(np.max(np.abs(o.expect[b, i] - np.real(r.expect[i]))) for i in range(len(r.expect)))
        print(f"{name} b={b}: gpu_def-truth td={d(outs['def'].final_state[b], truth.final_state):.2e} exp={de(outs['def'], truth):.2e} | "
              f"gpu_tight-truth td={d(outs['tight'].final_state[b], truth.final_state):.2e} exp={de(outs['tight'], truth):.2e} | "
              f"gpu_vtight-truth td={d(outs['vtight'].final_state[b], truth.final_state):.2e} | "
              f"orc_def-truth td={d(refs['def'].final_state, truth.final_state):.2e} | orc_tight-truth td={d(refs['tight'].final_state, truth.final_state):.2e} | "
              f"gpu_tight-orc_tight td={d(outs['tight'].final_state[b], refs['tight'].final_state):.2e} exp={de(outs['tight'], refs['tight']):.2e}", flush=True)
run("c2", w.c2_selective(2, 2), 2)
run("c3", w.c3_drag(2), 3)
run("c1", w.c1_rabi(5), 2)
run("c4", w.c4_two_transmons_bus(2), 1)
