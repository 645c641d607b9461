"""C5 (N = 1875) on one GPU: Hermitian-tile stepper against the full-matrix stepper of the large path - same steps, same
numbers, and the time of both.

    python tools/c5_step_check.py [n_ops=1] [cavity_levels=25] [modes=herm,full,herm]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencing_b200 import workloads  # noqa: E402
from sequencing_b200.solver import Engine, _lib  # noqa: E402


def main():
    n_ops = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    eng = Engine.get(0)
    res = {}
    modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["herm", "full", "herm"]
    for mode in modes:
        bp = workloads.c5_transmon_two_cavities(cavity_levels=levels, n_ops=n_ops)
        bp.method = _lib.METHOD_LARGE
        if mode == "full":
            os.environ["SEQ_LARGE_FULL"] = "1"
        try:
            out = eng.solve(bp)
        finally:
            os.environ.pop("SEQ_LARGE_FULL", None)
        torch.cuda.synchronize()
        res[mode] = out
        print(json.dumps({"mode": mode, "variant": out.variant, "N": bp.N, "n_t": int(bp.coeff.shape[-1]), "kernel_ms": out.kernel_ms,
                          "steps": int(out.stats[0, 0]), "rejected": int(out.stats[0, 1]), "rhs": int(out.stats[0, 2]),
                          "ms_per_step": out.kernel_ms / max(int(out.stats[0, 0]) + int(out.stats[0, 1]), 1), "launches": out.launches,
                          "trace": float(np.trace(out.final_state[0]).real), "status": int(out.status[0])}), flush=True)
    if "herm" not in res or "full" not in res:
        return
    a, f = res["herm"], res["full"]
    print(json.dumps({"max_abs_final": float(np.max(np.abs(a.final_state - f.final_state))), "max_abs_expect": float(np.max(np.abs(a.expect - f.expect))),
                      "stats_equal": bool(np.array_equal(a.stats[:, :3], f.stats[:, :3]))}), flush=True)


if __name__ == "__main__":
    main()
