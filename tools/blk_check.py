"""Block-local diagonal kernel (kernel_diagblk.cuh) against the plain diagonal kernel (SEQ_DIAG_BLK=0) and the generic FMA
kernel on one workload: outputs, step counts, kernel time; then launch shapes (SEQ_BLK_SHAPE=minb,kr,wreg) on the bench sweep.
    python tools/blk_check.py [c2|c2full|c4|c2bench] [shape ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
bp = {"c2": lambda: w.c2_selective(4, 37, sigma=25), "c2full": lambda: w.c2_selective(8, 37), "c4": lambda: w.c4_two_transmons_bus(296),
      "c2bench": lambda: w.c2_selective(32, 32), "c4bench": lambda: w.c4_two_transmons_bus(512)}[which]()
eng = Engine.get(0)


def run(tag, env, method=None):
    for k in ("SEQ_DIAG_BLK", "SEQ_BLK_MINB"):
        os.environ.pop(k, None)
    os.environ.update(env)
    bp.method = _lib.METHOD_AUTO if method is None else method
    eng.solve(bp)
    ms = []
    for _ in range(3):
        out = eng.solve(bp)
        ms.append(out.kernel_ms)
    print(f"{which} {tag}: {out.variant or out.method}\n   kernel_ms min {min(ms):.3f} -> {bp.B / min(ms) * 1e3:.0f} traj/s   steps acc {out.stats[:, 0].mean():.1f} rej "
          f"{out.stats[:, 1].mean():.1f} rhs {out.stats[:, 2].mean():.1f}  status!=0: {(out.status != 0).sum()}", flush=True)
    return out


res = {"blk": run("blk", {}), "plain": run("plain", {"SEQ_DIAG_BLK": "0"})}
if bp.B * bp.coeff.shape[-1] <= 200000:
    res["generic"] = run("generic", {}, _lib.METHOD_GENERIC)
a = res["blk"]
for k in ("plain", "generic"):
    if k in res:
        b = res[k]
        print("blk vs %s: max |d final| %.3g  max |d expect| %.3g  stats equal %s" % (k, np.abs(a.final_state - b.final_state).max(), np.abs(a.expect - b.expect).max(),
                                                                                      np.array_equal(a.stats[:, :3], b.stats[:, :3])))
for shape in sys.argv[2:]:
    run("blk minb " + shape, {"SEQ_BLK_MINB": shape})
