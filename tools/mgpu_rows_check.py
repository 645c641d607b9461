"""Row-sharded large-system solve across the ranks of a torchrun launch, checked on every rank against
the single-GPU large path (same kernels, no exchange).  Prints ROWS_SHARDED_OK from rank 0.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_rows_check.py [N] [n_t]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import random_problem, row_sparse_problem  # noqa: E402
from sequencing_b200.distributed import solve_rows_sharded  # noqa: E402
from sequencing_b200.solver import Engine, _lib  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ok = True
    # dense operators -> ZGEMM chain; diagonal-structured operators -> the structured RHS kernel on every rank's row block
    for n_c, n_ctd, structured in ((2, 1, False), (0, 0, False), (2, 1, True)):
        if structured:
            mk = lambda: row_sparse_problem(N=N, B=1, n_t=n_t, w_h=5, w_l=2, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N + n_c, hscale=0.5, store_states=True)
        else:
            mk = lambda: random_problem(N=N, B=1, n_t=n_t, n_ch=2, n_c=n_c, n_ctd=n_ctd, n_e=2, seed=N + n_c, hscale=0.3, store_states=True)
        eng = Engine.get(local)
        single = mk()
        single.method = _lib.METHOD_LARGE
        ref = eng.solve(single)
        hooks = solve_rows_sharded(mk(), engine=eng, native=False)      # torch.distributed hooks (a Python callback per collective)
        out = solve_rows_sharded(mk(), engine=eng)                      # the engine's own NCCL communicator
        # (two communicators may sum the error norm in different orders: equal to rounding, not bit for bit)
        same_paths = bool(np.max(np.abs(hooks.states - out.states)) < 1e-13 and np.max(np.abs(hooks.expect - out.expect)) < 1e-13 and np.array_equal(hooks.stats[:, :2], out.stats[:, :2]))
        d_states = float(np.max(np.abs(out.states - ref.states)))
        d_exp = float(np.max(np.abs(out.expect - ref.expect)))
        same_steps = bool(np.array_equal(out.stats[:, :2], ref.stats[:, :2]))
        good = (same_paths and bool(np.all(out.status == 0)) and d_states < 1e-12 and d_exp < 1e-12 and same_steps and out.method == "large"
                and ref.variant.startswith("large/diagonals" if structured else "large/zgemm"))
        print(f"[rank {rank}/{world}] N={N} n_c={n_c} n_ctd={n_ctd} {ref.variant}: states {d_states:.2e} expect {d_exp:.2e} steps_equal={same_steps} "
              f"collectives={getattr(out, 'comm_stats', None)} ok={good}", flush=True)
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0 and int(flag.item()) == 1:
        print("ROWS_SHARDED_OK", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
