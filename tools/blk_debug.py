"""diagblk vs generic on the random diagonal-structured problems of tests/test_gpu_parity.py (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import row_sparse_problem
from sequencing_b200.solver import Engine, _lib
from sequencing_b200 import workloads as w
eng = Engine.get(0)
cases = [(7, 3, 1, 2, 0), (12, 3, 1, 2, 1), (20, 5, 2, 2, 1), (31, 3, 1, 3, 0), (45, 3, 1, 4, 0), (46, 5, 1, 2, 0), (47, 3, 2, 2, 0), (48, 3, 1, 1, 1), (63, 5, 3, 1, 0),
         (64, 3, 1, 2, 0), (79, 3, 1, 2, 1), (90, 5, 1, 3, 0), (45, 3, 1, 0, 0)]
def cmp(tag, mk):
    a, g = mk(), mk()
    a.method, g.method = _lib.METHOD_DIAG, _lib.METHOD_GENERIC
    oa, og = eng.solve(a), eng.solve(g)
    print(f"{tag}: {oa.variant[:60]}.. status {oa.status.tolist()[:4]} dfinal {np.abs(oa.final_state - og.final_state).max():.2e} dexp {np.abs(oa.expect - og.expect).max():.2e} "
          f"stats eq {np.array_equal(oa.stats[:, :2], og.stats[:, :2])} {oa.stats[0, :3].tolist()} {og.stats[0, :3].tolist()}", flush=True)
for N, w_h, w_l, n_c, n_ctd in cases:
    cmp(f"N={N} wh={w_h} wl={w_l} nc={n_c} nctd={n_ctd}", lambda: row_sparse_problem(N=N, B=3, n_t=8, w_h=w_h, w_l=w_l, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N, hscale=0.3, store_states=True))
cmp("c4(4)", lambda: w.c4_two_transmons_bus(4))
cmp("c2 small cav4", lambda: w.c2_selective(2, 2, sigma=10, cavity_levels=4))
