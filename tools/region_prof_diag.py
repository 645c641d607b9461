"""Cycle breakdown of solve_diag_kernel's step loop (thread 0 of CTA 0, clock64 region timers).
Builds a separate profiling library (-DSEQ_PROFILE_REGIONS) next to the product one and runs one workload.

    python tools/region_prof_diag.py [c2|c2full|c4]      (on a GPU box)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

so = os.path.join(ROOT, "sequencing_b200", "_seq_engine_prof.so")
src = os.path.join(ROOT, "sequencing_b200", "csrc", "seq_engine.cu")
if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(os.path.join(os.path.dirname(src), f)) for f in os.listdir(os.path.dirname(src))):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DSEQ_PROFILE_REGIONS",
                           "-shared", "-Xcompiler", "-fPIC", "-o", so, src])
from sequencing_b200.solver import _lib
_lib.LIB_PATH = so
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine

which = sys.argv[1] if len(sys.argv) > 1 else "c2full"
eng = Engine.get(0)
if which == "c2worst":      # the 148 most expensive trajectories of the bench sweep, the most expensive one on CTA 0
    import copy
    full = w.c2_selective(32, 32)
    order = np.argsort(-eng.solve(full).stats[:, 2], kind="stable")[:148]
    bp = copy.copy(full)
    bp.state0, bp.coeff = full.state0[order], full.coeff[order]
    if full.coeff_scale is not None:
        bp.coeff_scale = full.coeff_scale[order]
else:
    bp = {"c2": lambda: w.c2_selective(4, 37, sigma=25), "c2full": lambda: w.c2_selective(4, 37), "c4": lambda: w.c4_two_transmons_bus(148)}[which]()
eng.solve(bp)
buf = (C.c_ulonglong * 16)()
eng.lib.seq_engine_debug_regions(buf, 1)
out = eng.solve(bp)
eng.lib.seq_engine_debug_regions(buf, 0)
cyc = np.array(list(buf), dtype=float)
rhs, steps = float(out.stats[0, 2]), float(out.stats[0, 0] + out.stats[0, 1])
names = {0: "commit + outputs (Tr(E rho), stores)", 1: "stage combine + Y write + G(t) build", 2: "RHS barrier", 3: "RHS terms (compute_K) + k store",
         4: "error estimate", 5: "error block sum (barrier)", 6: "controller (warp 0: step control, spline values) - call overhead", 7: "step barrier",
         8: "  controller: state load", 9: "  controller: ctl_update (pow)", 10: "  controller: output bookkeeping + ctl_next", 11: "  controller: spline values", 12: "  controller: state store"}
if "controller warp" in out.variant:
    names.update({5: "error partial sums handed to the controller warp", 6: "speculative outputs + first stage input of the next step", 7: "wait for the verdict",
                  8: "  controller warp: prediction of the next control block", 9: "  controller warp: waiting for the step (last stage + error sums)",
                  10: "  controller warp: block sum + verification", 11: "  controller warp: publish verdict", 12: "  -"})
print(out.variant)
tot = cyc[:8].sum() if "controller warp" in out.variant else cyc.sum()
d = out.stats[:, 3]
print("steps with err <= 2.9e-6:", float((d & 0xffff).mean()), " err <= 2.9e-7:", float((d >> 16).mean()), "of", steps)
print(f"{which}: method {out.method} kernel_ms {out.kernel_ms:.2f}  steps {steps:.0f} RHS {rhs:.0f}  cycles/step {tot / steps:.0f}  cycles/RHS {tot / rhs:.0f}")
print(f"CTA 0: {cyc[13] / steps:.0f} cycles/step in {cyc[15] / steps:.1f} ns/step = {cyc[13] / max(cyc[15], 1):.3f} GHz; longest CTA {cyc[14] / steps:.0f} cycles/step; "
      f"kernel {out.kernel_ms * 1e6 / steps:.1f} ns/step per wave of {-(-bp.B // 148)}")
for r in range(13):
    print(f"  {cyc[r] / steps:8.0f} cyc/step  {cyc[r] / rhs:8.0f} cyc/RHS  {cyc[r] / tot * 100:5.1f}%  {names[r]}")
