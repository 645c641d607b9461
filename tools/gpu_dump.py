"""GPU half of the error attribution (tools/gpu_attrib.py split in two so that the CPU oracle does not run on
GPU-box time): solve the attribution workloads at default / tight tolerances and save the outputs.
    on the GPU box:  python tools/gpu_dump.py gpurun_out/attrib_gpu.npz
    anywhere:        python tools/gpu_dump.py --compare gpurun_out/attrib_gpu.npz"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sequencing_b200 import workloads as w

CASES = {"c2": (lambda: w.c2_selective(2, 2), 4), "c3": (lambda: w.c3_drag(2), 3), "c1": (lambda: w.c1_rabi(5), 2),
         "c4": (lambda: w.c4_two_transmons_bus(2), 1)}
TOLS = {"def": (1e-8, 1e-6), "tight": (1e-10, 1e-8)}

if sys.argv[1] != "--compare":
    from sequencing_b200.solver import Engine
    eng = Engine.get(0)
    out = {}
    for name, (mk, nb) in CASES.items():
        bp = mk()
        for tag, (a, r) in TOLS.items():
            q = copy.copy(bp)
            q.atol, q.rtol = a, r
            o = eng.solve(q)
            assert np.all(o.status == 0)
            out[f"{name}_{tag}_final"] = o.final_state[:nb]
            out[f"{name}_{tag}_expect"] = o.expect[:nb]
            out[f"{name}_{tag}_stats"] = o.stats[:nb]
            print(name, tag, o.method, o.variant, o.stats[:nb, :3].tolist())
    np.savez(sys.argv[1], **out)
else:
    from helpers import oracle_solve, trace_distance
    z = np.load(sys.argv[2])
    for name, (mk, nb) in CASES.items():
        bp = mk()
        for b in range(nb):
            truth = oracle_solve(bp, b, atol=1e-13, rtol=1e-11, nsteps=100000)
            ref = oracle_solve(bp, b)
            line = f"{name} b={b}: oracle_def-truth td={trace_distance(ref.final_state, truth.final_state):.2e}"
            for tag in TOLS:
                fin, ex = z[f"{name}_{tag}_final"][b], z[f"{name}_{tag}_expect"][b]
                de = max(float(np.max(np.abs(ex[i] - np.real(truth.expect[i])))) for i in range(len(truth.expect)))
                line += f" | gpu_{tag}-truth td={trace_distance(fin, truth.final_state):.2e} exp={de:.2e}"
                if tag == "def":
                    line += f" gpu_def-oracle_def td={trace_distance(fin, ref.final_state):.2e}"
            print(line + f"  rhs={z[f'{name}_def_stats'][b, 2]}", flush=True)
