import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import row_sparse_problem
from sequencing_b200.solver import Engine, _lib
eng = Engine.get(0)
mk = lambda: row_sparse_problem(N=12, B=3, n_t=8, w_h=3, w_l=1, n_ch=2, n_c=0, n_ctd=1, n_e=2, seed=12, hscale=0.3, store_states=True)
g = mk(); g.method = _lib.METHOD_GENERIC
og = eng.solve(g)
for cap in ("0",):
    os.environ["SEQ_DIAG_SMEM_CAP"] = cap
    a = mk(); a.method = _lib.METHOD_DIAG
    oa = eng.solve(a)
    np.set_printoptions(precision=1, linewidth=200)
    print(oa.stats, og.stats)
    for b in range(3):
        ds = np.abs(oa.states[b] - og.states[b]).reshape(8, -1).max(axis=1)
        print("b", b, "per-output max diff", ds)
    b = int(np.argmax([np.abs(oa.states[b] - og.states[b]).max() for b in range(3)]))
    o = int(np.argmax(np.abs(oa.states[b] - og.states[b]).reshape(8, -1).max(axis=1) > 1e-13))
    print("first bad output", o, "of b", b)
    print((np.abs(oa.states[b, o] - og.states[b, o]) > 1e-13).astype(int))
