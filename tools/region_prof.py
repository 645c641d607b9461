"""Cycle breakdown of solve_dmma_kernel's step loop (thread 0 of CTA 0, clock64 region timers).
Builds a separate profiling library (-DSEQ_PROFILE_REGIONS) next to the product one and runs one workload.

    python tools/region_prof.py [c2|c2full]      (on a GPU box)
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

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
bp = w.c2_selective(4, 37, sigma=25) if which == "c2" else w.c2_selective(4, 37)
eng = Engine.get(0)
eng.solve(bp)
buf = (C.c_ulonglong * 16)()
eng.lib.seq_engine_debug_regions(buf, 1)
out = eng.solve(bp)
eng.lib.seq_engine_debug_regions(buf, 0)
cyc = np.array(list(buf), dtype=float)
rhs = float(out.stats[0, 2])
names = {11: "error estimate (stream reads + norm math)", 12: "error block sum", 13: "step control (pow)", 14: "outputs (Tr(E rho), stores)",
         15: "next step set-up (out_idx, ctl_next, spline values)", 10: "stage loop overhead", 0: "combine (k-stream reads)", 8: "barrier after combine",
         1: "put_Y + L0 prefetch issue + barrier", 2: "G(t) build + barrier", 3: "G*Y product + T write + barrier", 4: "A^dag read",
         5: "cp.async wait + barrier (per L)", 6: "L*Y product + T write + barrier (per L)", 7: "T*L^dag product (per L)", 9: "k-stream store"}
tot = cyc.sum()
print(f"{which}: kernel_ms {out.kernel_ms:.2f}  RHS evals (trajectory 0) {rhs:.0f}  cycles/RHS {tot / rhs:.0f}")
for r in (11, 12, 13, 14, 15, 10, 0, 8, 1, 2, 3, 4, 5, 6, 7, 9):
    print(f"  {cyc[r] / rhs:8.0f} cyc/RHS  {cyc[r] / tot * 100:5.1f}%  {names[r]}")
