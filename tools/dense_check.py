"""Tensor-pipe rate on DENSE operators (every fragment mask bit set): random problems at the bench shapes.
    python tools/dense_check.py   (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import random_problem
from sequencing_b200.solver import Engine
eng = Engine.get(0)
for N, B, n_c, n_t in ((45, 296, 4, 60), (48, 296, 4, 60), (90, 148, 9, 12), (96, 148, 4, 12), (200, 1, 4, 4), (512, 1, 4, 4)):
    bp = random_problem(N=N, B=B, n_t=n_t, n_ch=2, n_c=n_c, n_e=2, seed=N, hscale=0.3)
    for _ in range(2):
        out = eng.solve(bp)
    rhs = float(out.stats[:, 2].sum())
    tf = 8.0 * N ** 3 * (1 + 2 * n_c) * rhs / (out.kernel_ms * 1e-3) / 1e12
    print(f"dense N={N} B={B} n_c={n_c} method={out.method}: kernel {out.kernel_ms:.2f} ms, {rhs:.0f} RHS, {tf:.2f} algorithmic TFLOP/s = {tf / 37.0 * 100:.1f}% of DMMA peak", flush=True)
