"""One short solve per integration kernel / launch shape, for compute-sanitizer (memcheck, racecheck):
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
Kernels: solve_diagblk_kernel (128 x 2 CTAs/SM, 256, 512 with L2 streams), solve_diag_kernel (controller warp, plain, 2-CTA
cluster, streamed), solve_sparse_kernel, solve_dmma_kernel, solve_dmma_stream_kernel, solve_small_kernel, solve_generic_kernel
(density matrix and ket), the large path's structured RHS."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import random_problem, row_sparse_problem, scattered_sparse_problem
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine, _lib

eng = Engine.get(0)


def run(tag, bp, method=_lib.METHOD_AUTO, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    try:
        bp.method = method
        out = eng.solve(bp)
    finally:
        for k in (env or {}):
            del os.environ[k]
    print(f"{tag}: method {out.method} | {out.variant[:110]} | status ok {bool(np.all(out.status == 0))} steps {out.stats[:, 0].mean():.1f}", flush=True)


c2 = lambda: w.c2_selective(2, 2, sigma=5)                       # N=45, n_t=20
c4 = lambda: w.c4_two_transmons_bus(2)                           # N=90
run("diagblk 128x2", c2())
run("diagblk 512/L2", c4())
run("diagblk 256", w.c2_selective(1, 2, sigma=5, cavity_levels=30))      # N=90: R=30 -> 480?  (shape check only)
run("diag + controller warp", c2(), env={"SEQ_DIAG_BLK": "0"})
run("diag plain", c2(), env={"SEQ_DIAG_BLK": "0", "SEQ_DIAG_CW": "0"})
run("diag 2-CTA cluster", c2(), env={"SEQ_DIAG_BLK": "0", "SEQ_DIAG_CL": "2"})
run("diag streamed", c4(), env={"SEQ_DIAG_BLK": "0"})
run("sparse (ELL)", scattered_sparse_problem(N=40, B=2, n_t=6, n_ch=2, n_c=2, n_e=2, seed=3), _lib.METHOD_SPARSE)
run("dmma", random_problem(N=40, B=2, n_t=5, n_ch=2, n_c=2, n_e=2, seed=1), _lib.METHOD_DMMA)
run("dmma_stream", random_problem(N=70, B=1, n_t=4, n_ch=1, n_c=1, n_e=1, seed=2), _lib.METHOD_DMMA_STREAM)
run("small", w.c1_rabi(64))
run("generic dm", random_problem(N=12, B=2, n_t=6, n_ch=2, n_c=2, n_e=2, seed=4), _lib.METHOD_GENERIC)
run("generic ket", random_problem(N=20, B=2, n_t=6, n_ch=2, n_c=0, n_e=2, seed=5, ket=True), _lib.METHOD_GENERIC)
run("large/diagonals", row_sparse_problem(N=130, B=1, n_t=4, w_h=3, w_l=1, n_ch=2, n_c=2, n_e=1, seed=130, hscale=0.3))
