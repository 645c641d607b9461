import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import row_sparse_problem
from sequencing_b200.solver import Engine, _lib
from sequencing_b200 import Transmon, Cavity, System, get_sequence, workloads as w
eng = Engine.get(0)
def cmp(tag, mk):
    a, g = mk(), mk()
    a.method, g.method = _lib.METHOD_DIAG, _lib.METHOD_GENERIC
    oa, og = eng.solve(a), eng.solve(g)
    print(f"{tag}: {oa.variant[:50]}.. gath {oa.variant.split('gathers/element')[0][-4:] if 'gathers' in oa.variant else ''} dfinal {np.abs(oa.final_state - og.final_state).max():.2e} dexp {np.abs(oa.expect - og.expect).max():.2e} "
          f"stats eq {np.array_equal(oa.stats[:, :2], og.stats[:, :2])} {oa.stats[0, :3].tolist()} {og.stats[0, :3].tolist()}", flush=True)
for n_c, n_ctd, w_h, w_l in ((0, 0, 3, 1), (2, 0, 1, 1), (2, 0, 3, 1), (2, 1, 1, 1), (0, 1, 1, 1), (0, 1, 3, 1), (2, 1, 3, 1)):
    for N in (12, 20):
        cmp(f"N={N} nc={n_c} nctd={n_ctd} wh={w_h}", lambda: row_sparse_problem(N=N, B=3, n_t=8, w_h=w_h, w_l=w_l, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N, hscale=0.3, store_states=True))
