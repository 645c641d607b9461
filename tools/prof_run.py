"""One short solve of the C2-shaped workload for ncu captures (148 trajectories = one wave)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
eng = Engine.get(0)
if which == "c2":
    bp = w.c2_selective(4, 37, sigma=25)       # N=45, n_t=100, B=148
elif which == "c2bench":
    bp = w.c2_selective(32, 32)                # the bench launch itself: N=45, n_t=1000, B=1024
elif which == "c4bench":
    bp = w.c4_two_transmons_bus(512)
elif which == "c4":
    bp = w.c4_two_transmons_bus(148)
elif which == "c3":
    bp = w.c3_drag(2048)
elif which == "c1":
    bp = w.c1_rabi(16384)
for _ in range(2):
    out = eng.solve(bp)
print(which, "method", out.method, "kernel_ms", out.kernel_ms, "ok", bool(np.all(out.status == 0)), "steps", out.stats[:, 0].mean(),
      "rejected", out.stats[:, 1].mean(), "rhs_total", int(out.stats[:, 2].sum()), "B", bp.B)
