"""Per-trajectory work distribution of the C2 bench sweep (RHS evaluations, confirmed predictions) and kernel
time of its cheapest / most expensive 148 trajectories, with and without the controller warp."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine

bp = w.c2_selective(32, 32)
eng = Engine.get(0)
os.environ["SEQ_DIAG_CW"] = "1"
eng.solve(bp)
out = eng.solve(bp)
rhs, conf = out.stats[:, 2], out.stats[:, 3]
print("all 1024: kernel_ms %.2f rhs min/mean/max %d %.1f %d  confirmed min/mean/max %d %.1f %d" % (out.kernel_ms, rhs.min(), rhs.mean(), rhs.max(), conf.min(), conf.mean(), conf.max()))
print("rhs histogram:", np.histogram(rhs, bins=[3990, 4010, 4100, 4300, 4600, 5000, 5500, 6100])[0].tolist())
order = np.argsort(rhs, kind="stable")

def subset(idx):
    s = copy.copy(bp)
    s.state0 = bp.state0[idx]
    s.coeff = bp.coeff[idx] if bp.coeff.shape[0] > 1 else bp.coeff
    if bp.coeff_scale is not None:
        s.coeff_scale = bp.coeff_scale[idx]
    return s

for name, idx in (("cheapest 148", order[:148]), ("most expensive 148", order[-148:]), ("first 592 in order", np.arange(592)),
                  ("all 1024", np.arange(1024))):
    sb = subset(idx)
    for cw in ("1", "0"):
        os.environ["SEQ_DIAG_CW"] = cw
        eng.solve(sb)
        ms = min(eng.solve(sb).kernel_ms for _ in range(3))
        o = eng.solve(sb)
        print(f"{name}: CW={cw} kernel_ms {ms:.3f}  rhs mean {o.stats[:, 2].mean():.1f} max {o.stats[:, 2].max()}  confirmed mean {o.stats[:, 3].mean():.1f}")
