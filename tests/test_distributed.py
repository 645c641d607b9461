"""Batch sharding across ranks (world_size 2, gloo, CPU): slices, gather, and the deferred-sweep
path under torch.distributed.  The rank-local integration is done by the CPU oracle (tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from sequencing_b200.distributed import shard_bounds, shard_problem  # noqa: E402


def test_shard_bounds_cover_batch_exactly():
    for B in (0, 1, 2, 7, 51, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_problem_slices_per_trajectory_arrays_only():
    from helpers import random_problem
    bp = random_problem(N=3, B=7, n_t=6, n_ch=2, n_c=1, n_ctd=1, n_e=1, seed=0)
    parts = [shard_problem(bp, r, 3) for r in range(3)]
    assert [p.B for p in parts] == [3, 2, 2]
    np.testing.assert_array_equal(np.concatenate([p.state0 for p in parts]), bp.state0)
    np.testing.assert_array_equal(np.concatenate([p.coeff for p in parts]), bp.coeff)
    np.testing.assert_array_equal(np.concatenate([p.rate for p in parts]), bp.rate)
    assert all(p.Hk is bp.Hk and p.Lc is bp.Lc and p.E is bp.E for p in parts)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import OracleEngine, random_problem
        from sequencing_b200.distributed import solve_sharded
        from sequencing_b200.solver import engine as engine_mod

        fake = OracleEngine()
        engine_mod.Engine.get = classmethod(lambda cls, device_index=None: fake)
        bp = random_problem(N=3, B=5, n_t=8, n_ch=2, n_c=1, n_e=2, seed=3, hscale=0.2, store_states=True)
        full = solve_sharded(bp, engine=fake)
        ref = OracleEngine().solve(bp)
        ok = (np.allclose(full.expect, ref.expect, atol=0) and np.allclose(full.final_state, ref.final_state, atol=0)
              and np.allclose(full.states, ref.states, atol=0) and fake.batch_sizes == [3 if rank == 0 else 2])
        # the deferred sweep path shards too
        import sequencing_b200 as sq
        from sequencing_b200 import System, Transmon, get_sequence, deferred_solves

        qubit = Transmon("qubit", levels=2)
        system = System("system", modes=[qubit])
        res = []
        with deferred_solves():
            for amp in np.linspace(0.2, 1.0, 5):
                seq = get_sequence(system)
                with qubit.pulse_scale(amp):
                    qubit.rotate_x(np.pi)
                res.append(seq.run(qubit.fock_dm(0), e_ops=[qubit.fock_dm(1)]))
        pops = [float(r.expect[0][-1]) for r in res]
        ok = ok and fake.batch_sizes[-1] == (3 if rank == 0 else 2) and pops[-1] > 0.99 and pops[0] < 0.2
        q.put((rank, bool(ok), pops))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_and_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in got), got
    assert got[0][2] == got[1][2]     # every rank ends with the full sweep


def test_row_blocks_cover_rho_and_match_the_engine_granularity():
    """Row-sharded large-system mode (C5): blocks of m = 16*ceil(ceil(N/world)/16) rows, covering [0, N)."""
    from sequencing_b200.distributed import row_block
    for N in (97, 150, 243, 1875):
        for world in (1, 2, 4, 8):
            spans = [row_block(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            m = max(hi - lo for lo, hi in spans)
            assert m % 16 == 0 or world == 1 or m == spans[0][1] - spans[0][0]
            assert all((hi - lo) in (m,) or hi == N for lo, hi in spans)
    assert [row_block(1875, r, 8) for r in (0, 7)] == [(0, 240), (1680, 1875)]


def test_comm_struct_matches_header():
    """``struct seq_comm`` of include/seq_engine.h: two int32, a pointer, two function pointers."""
    import ctypes as C
    from sequencing_b200.solver import _lib
    assert [f[0] for f in _lib.Comm._fields_] == ["rank", "world", "ctx", "all_gather", "all_reduce_sum"]
    assert C.sizeof(_lib.Comm) == 32
    hdr = open(os.path.join(ROOT, "include", "seq_engine.h")).read()
    assert "typedef struct seq_comm" in hdr and "seq_engine_solve_sharded" in hdr
