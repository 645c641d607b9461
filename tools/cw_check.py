"""Controller-warp variant of solve_diag_kernel: how many predictions were confirmed (stats[:, 3]) and the kernel
time with / without the controller warp (SEQ_DIAG_CW=0) on one workload.   python tools/cw_check.py [c2|c2full|c4]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine

which = sys.argv[1] if len(sys.argv) > 1 else "c2full"
bp = {"c2": lambda: w.c2_selective(4, 37, sigma=25), "c2full": lambda: w.c2_selective(8, 37), "c4": lambda: w.c4_two_transmons_bus(296)}[which]()
eng = Engine.get(0)
res = {}
for cw in ("1", "0"):
    os.environ["SEQ_DIAG_CW"] = cw
    eng.solve(bp)
    out = eng.solve(bp)
    res[cw] = out
    print(f"{which} CW={cw}: {out.variant}\n   kernel_ms {out.kernel_ms:.3f}  steps acc {out.stats[:, 0].mean():.1f} rej {out.stats[:, 1].mean():.1f} rhs {out.stats[:, 2].mean():.1f}"
          f"  confirmed predictions {out.stats[:, 3].mean():.1f}")
a, b = res["1"], res["0"]
print("max |d final| %.3g  max |d expect| %.3g  stats equal %s" % (np.abs(a.final_state - b.final_state).max(), np.abs(a.expect - b.expect).max(),
                                                                     np.array_equal(a.stats[:, :3], b.stats[:, :3])))
