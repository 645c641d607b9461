"""Multi-GPU: batch sharding of sweeps (independent trajectories, no data-path collective) and the
row-sharded solve of ONE large system (an all-gather of the stage state per RHS evaluation).

Sweeps and initial-state batches are independent units: every rank integrates a
contiguous ``B/G`` slice of the batch on its own GPU with the shared operators replicated, and the
expectation traces / final states (/ stored states) are exchanged with ONE ``all_gather`` at the
end (NCCL over NVLink on GPUs; gloo on CPU in the tests).  There is no collective inside the
integration itself.  Launch with ``python -m torch.distributed.run --nproc-per-node G ...``.
"""
import copy
from typing import Optional

import numpy as np

from .solver.engine import BatchOutput, BatchProblem


def shard_bounds(B: int, rank: int, world: int):
    """Contiguous slice ``[lo, hi)`` of rank ``rank``; the first ``B % world`` ranks get one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_problem(bp: BatchProblem, rank: int, world: int) -> BatchProblem:
    """The rank-local part of ``bp``: per-trajectory arrays are sliced, shared operators kept."""
    lo, hi = shard_bounds(bp.B, rank, world)
    B = bp.B
    out = copy.copy(bp)

    def cut(a):
        if a is None:
            return None
        return a[lo:hi] if (a.ndim >= 1 and a.shape[0] == B and B > 1) else a

    out.state0 = bp.state0[lo:hi]
    out.coeff = cut(bp.coeff)
    out.rate = cut(bp.rate)
    out.coeff_scale = cut(bp.coeff_scale)
    if bp.H0.ndim == 3:
        out.H0 = cut(bp.H0)
    return out


def _all_gather_rows(local: np.ndarray, counts, device, dist, group):
    """all_gather of arrays that differ in their leading dimension (padded to the largest shard)."""
    import torch

    width = max(counts)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(np.ascontiguousarray(pad)).to(device)
    if t.is_complex():
        t = torch.view_as_real(t).contiguous()
    parts = [torch.empty_like(t) for _ in counts]
    dist.all_gather(parts, t, group=group)
    rows = []
    for n, part in zip(counts, parts):
        a = part.cpu().numpy()
        if np.iscomplexobj(local):
            a = a.view(np.complex128).reshape((width,) + local.shape[1:])
        rows.append(a[:n])
    return np.concatenate(rows, axis=0)


def solve_sharded(bp: BatchProblem, engine=None, group=None, gather: bool = True) -> BatchOutput:
    """Integrate ``bp`` across all ranks of the process group.

    Every rank passes the SAME full problem; each integrates its slice; with ``gather`` the full
    ``BatchOutput`` is returned on every rank (otherwise only the local shard's output).
    """
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        from .solver import Engine

        return (engine or Engine.get()).solve(bp)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        from .solver import Engine

        engine = Engine.get(torch.cuda.current_device())
    local = shard_problem(bp, rank, world)
    counts = [hi - lo for lo, hi in (shard_bounds(bp.B, r, world) for r in range(world))]
    n_e = bp.E.shape[0] if bp.E is not None else 0
    if local.B > 0:
        out = engine.solve(local)
    else:  # more ranks than trajectories: contribute an empty shard
        st = (bp.N,) if bp.is_ket else (bp.N, bp.N)
        n_out = len(bp.out_idx)
        out = BatchOutput(expect=np.zeros((0, n_e, n_out), dtype=complex if bp.expect_complex else float),
                          final_state=np.zeros((0,) + st, complex),
                          states=np.zeros((0, n_out) + st, complex) if bp.store_states else None,
                          status=np.zeros(0, np.int32), stats=np.zeros((0, 4), np.int32))
    if not gather:
        return out
    backend = dist.get_backend(group)
    if backend == "nccl":
        import torch

        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = "cpu"
    g = lambda a: _all_gather_rows(a, counts, device, dist, group)
    return BatchOutput(
        expect=g(out.expect), final_state=g(out.final_state),
        states=g(out.states) if out.states is not None else None,
        status=g(out.status), stats=g(out.stats), kernel_ms=out.kernel_ms, launches=out.launches, method=out.method, variant=getattr(out, "variant", ""),
    )


# ---------------------------------------------------------------------------------------------
# Row-sharded solve of one large system (C5): rho split into row blocks, one all-gather per RHS
# ---------------------------------------------------------------------------------------------
def row_block(N: int, rank: int, world: int):
    """Rows ``[lo, hi)`` of rho that rank ``rank`` evaluates: blocks of m = 16*ceil(ceil(N/world)/16)
    rows (the engine's tile granularity), the tail ranks may be short or empty."""
    m = ((N + world - 1) // world + 15) // 16 * 16
    lo = min(N, rank * m)
    return lo, min(N, lo + m)


class _DevArray:
    """A device pointer as an object torch can wrap without copying (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_comm(group=None, stream=None):
    """``seq_comm`` hooks backed by torch.distributed (NCCL): the engine calls them with device
    pointers while it enqueues the stepper on ITS stream (the one captured when the ``Engine`` was
    created), so the hooks issue the collectives under ``torch.cuda.stream(stream)`` - ordered with the
    engine's kernels whatever torch's current stream is at call time.  ``stream``: the engine's
    ``torch.cuda.Stream`` (``Engine.stream``); None = torch's current stream.
    Returns (comm, keepalive) - keep ``keepalive`` referenced during the solve."""
    import contextlib

    import torch
    import torch.distributed as dist

    from .solver import _lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    stats = {"all_gather": 0, "all_reduce": 0, "bytes_gathered": 0}

    def _wrap(ptr, n):
        return torch.as_tensor(_DevArray(ptr, n), device=torch.device("cuda", torch.cuda.current_device()))

    def _on_stream():
        return torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()

    def all_gather(ctx, send, recv, count):
        try:
            with _on_stream():
                dist.all_gather_into_tensor(_wrap(recv, count * world), _wrap(send, count), group=group)
            stats["all_gather"] += 1
            stats["bytes_gathered"] += 8 * count * world
            return 0
        except Exception as exc:  # the C side turns a non-zero return into an error message
            print(f"[seq_comm] all_gather failed: {exc}", flush=True)
            return 1

    def all_reduce(ctx, buf, count):
        try:
            with _on_stream():
                dist.all_reduce(_wrap(buf, count), group=group)
            stats["all_reduce"] += 1
            return 0
        except Exception as exc:
            print(f"[seq_comm] all_reduce failed: {exc}", flush=True)
            return 1

    ag, ar = _lib.ALL_GATHER_FN(all_gather), _lib.ALL_REDUCE_FN(all_reduce)
    comm = _lib.Comm(rank=rank, world=world, ctx=None, all_gather=ag, all_reduce_sum=ar)
    return comm, (ag, ar, stats)


def solve_rows_sharded(bp: BatchProblem, engine=None, group=None, native=True) -> BatchOutput:
    """Integrate ONE large Lindblad system with rho row-sharded over every rank of the process group
    (config C5: transmon + two cavities, N = 1875, 8 x B200).  Every rank passes the same problem and
    receives the full result.  Without an initialised process group this is a single-GPU solve."""
    import torch
    import torch.distributed as dist

    from .solver import Engine

    engine = engine or Engine.get(torch.cuda.current_device())
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return engine.solve_rows_sharded(bp, None)
    if native:
        # NCCL calls issued by the engine itself (C, its own stream): no Python callback per collective
        if getattr(engine, "native_world", 0) != dist.get_world_size(group):
            engine.init_native_comm(group)
        out = engine.solve_rows_sharded(bp, None)
        out.comm_stats = {"native_nccl": True}
        return out
    comm, keep = make_comm(group, stream=engine.stream)
    out = engine.solve_rows_sharded(bp, comm)
    out.comm_stats = keep[2]
    return out
