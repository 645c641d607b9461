import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine
eng = Engine.get(0)
for name, bp in (("c2", w.c2_selective(32, 32)), ("c3", w.c3_drag(4096)), ("c1", w.c1_rabi(65536))):
    for label, fn in (("solve (torch staging)", lambda: eng.solve(bp)), ("solve_host_abi pageable out", lambda: eng.solve_host_abi(bp)),
                      ("solve_host_abi pinned out", lambda: eng.solve_host_abi(bp, pinned_out=True))):
        fn(); fn()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); o = fn(); ts.append(time.perf_counter() - t0)
        print(f"{name} B={bp.B} {label}: {min(ts)*1e3:.1f} ms  ({bp.B/min(ts):.0f} traj/s) method {o.method}", flush=True)
