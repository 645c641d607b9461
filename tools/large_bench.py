"""Time the large-system path (SEQ_METHOD_LARGE) on the C5 shape: one GPU, or row-sharded under torchrun.

    python tools/large_bench.py [cavity_levels=25] [n_ops=1] [out_samples=8]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/large_bench.py 25 1 8

Only the first `out_samples` samples of the sequence are integrated (the full C5 sequence is 10^4 samples);
prints one JSON line with the algorithmic FP64 rate 8 N^3 (1 + 2 n_c) per RHS evaluation.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencing_b200 import workloads  # noqa: E402
from sequencing_b200.distributed import solve_rows_sharded  # noqa: E402
from sequencing_b200.solver import Engine, _lib  # noqa: E402


def main():
    levels = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    n_ops = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bp = workloads.c5_transmon_two_cavities(cavity_levels=levels, n_ops=n_ops)
    bp.out_idx = np.arange(min(n_out, bp.coeff.shape[-1]), dtype=np.int32)
    bp.method = _lib.METHOD_LARGE
    eng = Engine.get(local)
    res = []
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = solve_rows_sharded(bp, engine=eng)
        torch.cuda.synchronize()
        res.append(time.perf_counter() - t0)
    n_c = bp.Lc.shape[0]
    rhs = int(out.stats[0, 2])
    flops = 8.0 * bp.N ** 3 * (1 + 2 * n_c) * rhs
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"N": bp.N, "world": world, "n_c": int(n_c), "n_ch": int(bp.Hk.shape[0]), "steps": int(out.stats[0, 0]),
                          "rejected": int(out.stats[0, 1]), "rhs_evals": rhs, "kernel_ms": out.kernel_ms, "wall_s": res,
                          "ms_per_rhs": out.kernel_ms / max(rhs, 1), "algorithmic_TFLOPs": flops / (out.kernel_ms * 1e-3) / 1e12,
                          "status_ok": bool(np.all(out.status == 0)), "trace": float(np.trace(out.final_state[0]).real),
                          "expect_last": out.expect[0, :, -1].tolist(), "launches": out.launches,
                          "comm": getattr(out, "comm_stats", None)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
