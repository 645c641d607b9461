#!/usr/bin/env python
"""Regenerate the golden fixtures under tests/golden/ (run from the repo root on CPU):

    python tests/golden/make_golden.py

The reference path bottoms out in QuTiP 4.x, which is neither under /root/reference nor
installable here, and the reference ships no stored traces, so these vectors are produced by the
CPU oracle (oracle/qutip4_oracle.py - the QuTiP-4.x restatement on scipy's ZVODE) on inputs lowered
through the public host API (sequencing_b200.workloads: System / Transmon / Cavity / seq.run).

Each ``<case>.npz`` holds, for the first ``n`` trajectories of the named workload:
  inputs : H0, Hk, coeff, Lc, Ltd, rate, state0, E, out_idx, t0, dt, max_step       (what the engine receives)
  oracle : expect_default / final_default   ZVODE-Adams at the reference defaults (atol 1e-8, rtol 1e-6)
           expect_tight   / final_tight     ZVODE-Adams at atol 1e-10, rtol 1e-8
           expect_truth   / final_truth     converged run (DOP853 rtol 1e-13, or ZVODE 1e-13/1e-11 for stiff cases)
tests/test_golden.py checks (CPU) that the oracle still reproduces them and (GPU) that the CUDA engine
matches them within the north-star tolerance.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import oracle_solve, random_problem  # noqa: E402
from sequencing_b200 import workloads  # noqa: E402


def cases():
    """name -> (builder, number of trajectories stored, truth via DOP853?)"""
    return {
        "c1_rabi": (lambda: workloads.c1_rabi(9), 9, True),
        "c3_drag_xpy9": (lambda: workloads.c3_drag(2), 12, True),
        "c3_drag_ypx9_lossless": (lambda: workloads.c3_drag(2, family="YpX9", with_loss=False), 6, True),
        "c2_small": (lambda: workloads.c2_selective(2, 2, sigma=10, cavity_levels=4), 4, True),
        "c2_full": (lambda: workloads.c2_selective(2, 2), 2, False),
        "c4_bus": (lambda: workloads.c4_two_transmons_bus(2), 1, False),
        "random_n5_ctd": (lambda: random_problem(N=5, B=3, n_t=40, n_ch=2, n_c=1, n_ctd=2, n_e=2, seed=5, hscale=0.3), 3, True),
        "random_ket_n7": (lambda: random_problem(N=7, B=3, n_t=30, n_ch=2, n_c=0, n_e=2, seed=17, ket=True, hscale=0.3), 3, True),
    }


def take(bp, n):
    """First n trajectories of a BatchProblem (inputs only)."""
    import copy

    out = copy.copy(bp)
    out.state0 = bp.state0[:n]
    if bp.coeff is not None and bp.coeff.shape[0] > 1:
        out.coeff = bp.coeff[:n]
    if bp.rate is not None and bp.rate.shape[0] > 1:
        out.rate = bp.rate[:n]
    if bp.coeff_scale is not None:
        out.coeff_scale = bp.coeff_scale[:n]
    return out


def solve_all(bp, dop853):
    n_e = bp.E.shape[0] if bp.E is not None else 0
    res = {}
    for tag in ("default", "tight", "truth"):
        ex, fin = [], []
        for b in range(bp.B):
            if tag == "default":
                r = oracle_solve(bp, b)
            elif tag == "tight":
                r = oracle_solve(bp, b, atol=1e-10, rtol=1e-8, nsteps=100000)
            elif dop853:
                r = oracle_solve(bp, b, truth=True)
            else:
                r = oracle_solve(bp, b, atol=1e-13, rtol=1e-11, nsteps=1000000)
            ex.append(np.stack([np.asarray(r.expect[m]) for m in range(n_e)]) if n_e else np.zeros((0, len(bp.out_idx))))
            fin.append(np.asarray(r.final_state).reshape(bp.state0.shape[1:]))
        res[f"expect_{tag}"] = np.stack(ex)
        res[f"final_{tag}"] = np.stack(fin)
    return res


def main(only=None):
    for name, (make, n, dop853) in cases().items():
        if only and name not in only:
            continue
        bp = take(make(), n)
        data = dict(
            H0=bp.H0, Hk=bp.Hk, coeff=bp.coeff, Lc=bp.Lc, Ltd=bp.Ltd, rate=bp.rate, state0=bp.state0, E=bp.E,
            out_idx=bp.out_idx, t0=bp.t0, dt=bp.dt, max_step=bp.max_step, is_ket=bp.is_ket, truth_is_dop853=dop853,
        )
        data.update(solve_all(bp, dop853))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: B={bp.B} N={bp.N} n_t={bp.coeff.shape[-1]} -> {os.path.getsize(path) / 1024:.1f} KiB", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
