"""Shared test helpers: oracle adapters and canonical synthetic problems."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import qutip4_oracle as orc  # noqa: E402
from sequencing_b200.solver.engine import BatchOutput, BatchProblem  # noqa: E402


def oracle_options(bp: BatchProblem, **over):
    kw = dict(atol=bp.atol, rtol=bp.rtol, nsteps=bp.nsteps, first_step=bp.first_step, min_step=bp.min_step,
              max_step=bp.max_step, store_states=bp.store_states, normalize_output=bp.normalize)
    kw.update(over)
    return orc.Options(**kw)


def oracle_inputs(bp: BatchProblem, b: int):
    """Trajectory ``b`` of a BatchProblem in the oracle's list format."""
    n_t = bp.coeff.shape[-1] if bp.coeff is not None and bp.coeff.size else (bp.rate.shape[-1] if bp.rate is not None and bp.rate.size else int(bp.out_idx.max()) + 1)
    grid = bp.t0 + bp.dt * np.arange(n_t)
    H0 = bp.H0[b] if bp.H0.ndim == 3 else bp.H0
    H = [H0]
    n_ch = bp.Hk.shape[0] if bp.Hk is not None else 0
    for k in range(n_ch):
        c = bp.coeff[b if bp.coeff.shape[0] > 1 else 0, k]
        if bp.coeff_scale is not None:
            c = c * bp.coeff_scale[b, k]
        H.append((bp.Hk[k], c))
    c_ops = [L for L in (bp.Lc if bp.Lc is not None else [])]
    n_ctd = bp.Ltd.shape[0] if bp.Ltd is not None else 0
    for j in range(n_ctd):
        r = bp.rate[b if bp.rate.shape[0] > 1 else 0, j]
        c_ops.append((bp.Ltd[j], np.sqrt(r)))
    e_ops = [E for E in (bp.E if bp.E is not None else [])]
    tlist = grid[bp.out_idx]
    return H, bp.state0[b], tlist, c_ops, e_ops, grid


def oracle_solve(bp: BatchProblem, b: int, truth=False, **opt_over):
    H, s0, tlist, c_ops, e_ops, grid = oracle_inputs(bp, b)
    if truth:
        return orc.truth_solve(H, s0, tlist, c_ops, e_ops, coeff_tlist=grid)
    return orc.mesolve(H, s0, tlist, c_ops, e_ops, oracle_options(bp, **opt_over), coeff_tlist=grid)


class OracleEngine:
    """Engine look-alike that integrates each trajectory with the CPU oracle (tests only)."""

    def __init__(self):
        self.calls = 0
        self.batch_sizes = []

    def solve(self, bp: BatchProblem) -> BatchOutput:
        self.calls += 1
        B, N = bp.B, bp.N
        self.batch_sizes.append(B)
        n_e = bp.E.shape[0] if bp.E is not None else 0
        n_out = len(bp.out_idx)
        st = (N,) if bp.is_ket else (N, N)
        expect = np.zeros((B, n_e, n_out), dtype=complex if bp.expect_complex else float)
        final = np.zeros((B,) + st, dtype=complex)
        states = np.zeros((B, n_out) + st, dtype=complex) if bp.store_states else None
        for b in range(B):
            r = oracle_solve(bp, b, store_states=True)
            for m in range(n_e):
                expect[b, m] = r.expect[m] if bp.expect_complex else np.real(r.expect[m])
            final[b] = r.final_state.reshape(st)
            if states is not None:
                states[b] = np.stack([s.reshape(st) for s in r.states])
        return BatchOutput(expect=expect, final_state=final, states=states, status=np.zeros(B, np.int32),
                           stats=np.zeros((B, 4), np.int32), method="oracle")

    solve_host_abi = solve


def trace_distance(a, b):
    return orc.trace_distance(a, b)


# ---- canonical synthetic problems ----------------------------------------------------------
def random_problem(N, B, n_t, n_ch=2, n_c=2, n_ctd=0, n_e=2, seed=0, ket=False, dt=1.0, hscale=0.5, **kw):
    """Seeded random dense problem (Hermitian H terms, random collapse ops, smooth coefficients)."""
    rng = np.random.default_rng(seed)

    def herm():
        a = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
        return hscale * (a + a.conj().T) / (2 * np.sqrt(N))

    H0 = herm()
    Hk = np.stack([herm() for _ in range(n_ch)]) if n_ch else np.zeros((0, N, N), complex)
    ts = np.arange(n_t) * dt
    coeff = np.zeros((B, n_ch, n_t))
    for b in range(B):
        for k in range(n_ch):
            f = rng.uniform(0.02, 0.08) / dt
            coeff[b, k] = rng.uniform(0.5, 1.5) * np.sin(2 * np.pi * f * ts + rng.uniform(0, 6.28)) * np.exp(-((ts - ts.mean()) / (0.35 * max(ts[-1], dt))) ** 2)
    Lc = np.stack([0.15 * (rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))) / np.sqrt(N) for _ in range(n_c)]) if n_c else np.zeros((0, N, N), complex)
    Ltd = np.stack([0.15 * (rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))) / np.sqrt(N) for _ in range(n_ctd)]) if n_ctd else np.zeros((0, N, N), complex)
    rate = np.zeros((B, n_ctd, n_t))
    for b in range(B):
        for j in range(n_ctd):
            rate[b, j] = rng.uniform(0.2, 1.0) * (1 + np.cos(2 * np.pi * ts / max(ts[-1], dt) + rng.uniform(0, 6.28))) / 2
    if ket:
        psi = rng.normal(size=(B, N)) + 1j * rng.normal(size=(B, N))
        state0 = psi / np.linalg.norm(psi, axis=1, keepdims=True)
    else:
        state0 = np.zeros((B, N, N), complex)
        for b in range(B):
            a = rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
            rho = a @ a.conj().T
            state0[b] = rho / np.trace(rho).real
    E = np.stack([herm() for _ in range(n_e)]) if n_e else np.zeros((0, N, N), complex)
    args = dict(H0=H0, Hk=Hk, coeff=coeff, state0=state0, t0=0.0, dt=dt, out_idx=np.arange(n_t, dtype=np.int32),
                Lc=Lc, Ltd=Ltd, rate=rate, E=E, is_ket=ket, max_step=dt)
    args.update(kw)
    return BatchProblem(**args)


def row_sparse_problem(N, B, n_t, w_h=3, w_l=1, seed=0, **kw):
    """``random_problem`` with row-sparse operators: every Hamiltonian term keeps ``w_h`` random
    (Hermitian-symmetrised) off-diagonals plus the diagonal, every collapse operator keeps ``w_l``
    entries per row on shifted diagonals with wrap-around holes - the structure of the reference's
    Kronecker-embedded ladder / number operators (modes.py:125-207), at random values."""
    bp = random_problem(N=N, B=B, n_t=n_t, seed=seed, **kw)
    rng = np.random.default_rng(1000 + seed)
    i, j = np.indices((N, N))

    def band_mask(offsets):
        m = np.zeros((N, N), bool)
        for o in offsets:
            m |= (j - i) == o
        return m

    offs = [int(o) for o in rng.choice(np.arange(1, max(2, N)), size=min(max(w_h // 2, 0), N - 1), replace=False)]
    hmask = band_mask([0] + offs + [-o for o in offs])
    bp.H0 = bp.H0 * hmask * np.sqrt(N)
    bp.Hk = bp.Hk * hmask * np.sqrt(N)
    if bp.H0.ndim == 2 and w_h % 2 == 0:
        bp.H0 = bp.H0 * (1 - np.eye(N)) + np.diag(np.zeros(N))     # even width: no diagonal

    def lmask():
        lo = [int(o) for o in rng.choice(np.arange(-(N - 1), N), size=w_l, replace=False)]
        m = band_mask(lo)
        m &= rng.random((N, N)) < 0.8      # some rows lose entries (ragged ELL rows)
        return m

    if bp.Lc is not None and bp.Lc.shape[0]:
        bp.Lc = np.stack([L * lmask() * np.sqrt(N) for L in bp.Lc])
    if bp.Ltd is not None and bp.Ltd.shape[0]:
        bp.Ltd = np.stack([L * lmask() * np.sqrt(N) for L in bp.Ltd])
    if bp.E is not None and bp.E.shape[0] > 1:
        v = np.zeros(N); v[seed % N] = 1.0
        bp.E[0] = np.outer(v, v)           # a projector (1 non-zero), as the sweeps use
    return bp


def scattered_sparse_problem(N, B, n_t, seed=0, **kw):
    """Row-sparse operators WITHOUT diagonal structure: every Hamiltonian term is P + P^dag for a random
    weighted permutation P plus a diagonal (<= 3 non-zeros per row on ~N different diagonals), every
    collapse operator a random weighted permutation (1 non-zero per row)."""
    bp = random_problem(N=N, B=B, n_t=n_t, seed=seed, **kw)
    rng = np.random.default_rng(2000 + seed)

    def perm():
        P = np.zeros((N, N), complex)
        P[np.arange(N), rng.permutation(N)] = rng.normal(size=N) + 1j * rng.normal(size=N)
        return P

    def herm():
        P = perm()
        return 0.3 * (P + P.conj().T) + np.diag(rng.normal(size=N))

    bp.H0 = herm()
    bp.Hk = np.stack([herm() for _ in bp.Hk]) if bp.Hk.shape[0] else bp.Hk
    if bp.Lc is not None and bp.Lc.shape[0]:
        bp.Lc = np.stack([0.2 * perm() for _ in bp.Lc])
    if bp.Ltd is not None and bp.Ltd.shape[0]:
        bp.Ltd = np.stack([0.2 * perm() for _ in bp.Ltd])
    return bp


# ---- sampled parity against the oracle (tests/test_parity_sampled.py and bench.py's `parity` block) ----------
def take_one(bp: BatchProblem, b: int) -> BatchProblem:
    """Trajectory ``b`` of a BatchProblem as a one-trajectory problem (small enough to pickle to a worker)."""
    import copy

    out = copy.copy(bp)
    out.state0 = np.ascontiguousarray(bp.state0[b:b + 1])
    if bp.coeff is not None and bp.coeff.shape[0] > 1:
        out.coeff = np.ascontiguousarray(bp.coeff[b:b + 1])
    if bp.rate is not None and bp.rate.shape[0] > 1:
        out.rate = np.ascontiguousarray(bp.rate[b:b + 1])
    if bp.coeff_scale is not None:
        out.coeff_scale = np.ascontiguousarray(bp.coeff_scale[b:b + 1])
    if bp.H0.ndim == 3 and bp.H0.shape[0] > 1:
        out.H0 = np.ascontiguousarray(bp.H0[b])
    return out


def _oracle_task(args):
    """Worker: one trajectory through the oracle at (atol, rtol).  Returns (final_state, expect[n_e, n_out])."""
    one, atol, rtol = args
    r = oracle_solve(one, 0, atol=atol, rtol=rtol, nsteps=max(int(one.nsteps), 1000000 if atol < 1e-9 else 1000))
    n_e = one.E.shape[0] if one.E is not None else 0
    ex = np.stack([np.asarray(r.expect[m]) for m in range(n_e)]) if n_e else np.zeros((0, len(one.out_idx)))
    return np.asarray(r.final_state), ex


def oracle_many(bp: BatchProblem, idx, tolerances, workers=None):
    """``{(atol, rtol): [(final_state, expect), ...]}`` for trajectories ``idx`` of ``bp``, computed by a pool of
    SPAWNED worker processes (the parent may hold a CUDA context; the workers never touch CUDA)."""
    import concurrent.futures as cf
    import multiprocessing as mp

    tasks = [(take_one(bp, int(b)), a, r) for (a, r) in tolerances for b in idx]
    workers = workers or min(len(tasks), os.cpu_count() or 1)
    if workers <= 1:
        res = [_oracle_task(t) for t in tasks]
    else:
        with cf.ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_oracle_task, tasks, chunksize=1))
    out, k = {}, 0
    for tol in tolerances:
        out[tuple(tol)] = res[k:k + len(idx)]
        k += len(idx)
    return out


def _fraction_report(idx, get_a, get_b, tol):
    """Outputs (one final state + n_e expectation traces per trajectory) of a within ``tol`` of b."""
    n_ok = n_all = n_traj_ok = 0
    worst_td = worst_ee = 0.0
    for q, b in enumerate(idx):
        fa, ea = get_a(q, b)
        fb, eb = get_b(q, b)
        td = trace_distance(fa, np.asarray(fb).reshape(np.asarray(fa).shape))
        ok_traj = td <= tol
        n_ok += int(td <= tol); n_all += 1
        worst_td = max(worst_td, td)
        for m in range(len(eb)):
            ee = float(np.max(np.abs(np.asarray(ea[m]) - np.asarray(eb[m])))) if np.iscomplexobj(ea) else float(np.max(np.abs(np.real(ea[m]) - np.real(eb[m]))))
            n_ok += int(ee <= tol); n_all += 1
            ok_traj = ok_traj and ee <= tol
            worst_ee = max(worst_ee, ee)
        n_traj_ok += int(ok_traj)
    return {"outputs_within_tol": n_ok, "outputs": n_all, "fraction": n_ok / max(n_all, 1),
            "trajectories_within_tol": n_traj_ok, "trajectories": len(idx),
            "worst_trace_distance": worst_td, "worst_expect_abs_err": worst_ee, "tol": tol}


def parity_fractions(bp: BatchProblem, idx, engine_outputs, oracle_outputs, tol=1e-6, converged=None):
    """Per tolerance pair: the fraction of sampled OUTPUTS (one final state + n_e expectation traces per
    trajectory) of the engine within ``tol`` of the oracle AT THE SAME atol/rtol - final state by trace distance
    (kets: projector distance), expectation traces by max-abs error - plus the worst values.  ``converged``: the
    tolerance pair whose oracle run is converged (< 1e-7 from the exact solution); every other pair then also gets
    ``engine_vs_converged`` and ``oracle_vs_converged`` - who is how far from the exact solution at that pair."""
    rep = {}
    for key, refs in oracle_outputs.items():
        out = engine_outputs[key]
        eng = lambda q, b, out=out: (out.final_state[b], out.expect[b])
        orc_ = lambda q, b, refs=refs: refs[q]
        r = _fraction_report(idx, eng, orc_, tol)
        if converged is not None and tuple(converged) != tuple(key):
            conv = oracle_outputs[tuple(converged)]
            cv = lambda q, b, conv=conv: conv[q]
            r["engine_vs_converged"] = _fraction_report(idx, eng, cv, tol)
            r["oracle_vs_converged"] = _fraction_report(idx, orc_, cv, tol)
        rep["atol=%g,rtol=%g" % key] = r
    return rep
