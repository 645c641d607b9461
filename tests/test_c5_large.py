"""C5 (transmon(3) x cavity(25) x cavity(25), N = 1875, ONE system): the large path of the engine against a CPU truth
built from N x N SPARSE operators acting on the dense 1875 x 1875 rho - no N^2 x N^2 Liouvillian, no code shared with
the engine or the oracle (scipy.sparse products, scipy's natural CubicSpline, scipy's DOP853 at rtol 1e-10).
SURVEY.md 8(d) C5 / 8(e) row 2; the reference call is ``CompiledPulseSequence.run`` (basic.py:504-512) on the whole
sequence.  The CPU truth costs 0.6 s per right-hand side at N = 1875 (13 sparse-times-dense products of 56 MB), so the
full-size check covers the sequence's first operation (a cavity displacement, n_t = 12) and the same large path is
checked on the first four operations (displacements of both cavities and a qubit rotation, n_t = 48) of the 3 x 10 x 10
system (N = 300)."""
import time

import numpy as np
import pytest


def sparse_truth(bp, rtol=1e-10, atol=1e-12):
    """rho(t) at every grid point by DOP853 on  rho' = A + A^dag + sum_l L rho L^dag,  A = -i G(t) rho,
    G(t) = H0 + sum_k c_k(t) H_k - (i/2) sum_l L^dag L  with sparse N x N operators (scipy.sparse for the algebra, torch's
    multi-threaded CPU sparse-times-dense products for the 13 products per right-hand side)."""
    import warnings

    import scipy.sparse as sp
    import torch
    from scipy.integrate import solve_ivp
    from scipy.interpolate import CubicSpline

    warnings.filterwarnings("ignore", message=".*[Ss]parse.*")
    N = bp.N

    use_torch = N > 400      # (small systems: scipy's single-threaded products are faster than torch's dispatch)

    def to_t(m):
        m = sp.csr_matrix(m).astype(complex)
        m.sort_indices()
        if not use_torch:
            return m
        return torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int64)), torch.from_numpy(m.data), size=(N, N))

    H0 = sp.csr_matrix(bp.H0)
    Hk = [sp.csr_matrix(h) for h in bp.Hk]
    Ls = [sp.csr_matrix(L) for L in bp.Lc]
    assert bp.Ltd is None or bp.Ltd.shape[0] == 0
    G0 = (H0 - 0.5j * sum((L.conj().T @ L) for L in Ls)).tocsr()
    Lt = [to_t(L) for L in Ls]
    n_t = bp.coeff.shape[-1]
    grid = bp.t0 + bp.dt * np.arange(n_t)
    splines = [CubicSpline(grid, bp.coeff[0, k], bc_type="natural") for k in range(len(Hk))]
    nrhs = [0]

    def rhs(t, v):
        nrhs[0] += 1
        rho = torch.from_numpy(v).reshape(N, N) if use_torch else v.reshape(N, N)
        G = G0
        for k, h in enumerate(Hk):
            G = G + float(splines[k](t)) * h
        A = -1j * (to_t(G) @ rho)
        out = A + A.conj().T
        for L in Lt:
            X = (L @ rho).conj().T                 # (L rho)^dag
            out += (L @ (X.contiguous() if use_torch else np.ascontiguousarray(X))).conj().T      # (L rho) L^dag = (L (L rho)^dag)^dag
        return out.reshape(-1).numpy() if use_torch else out.reshape(-1)

    sol = solve_ivp(rhs, (grid[0], grid[-1]), np.ascontiguousarray(bp.state0[0].astype(complex).reshape(-1)), method="DOP853", t_eval=grid[bp.out_idx], rtol=rtol, atol=atol)
    assert sol.success
    states = sol.y.T.reshape(len(bp.out_idx), N, N)
    expect = np.array([[np.sum(E.T * r).real for r in states] for E in bp.E])      # Tr(E rho)
    return states[-1], expect, nrhs[0]


def _trace_distance(a, b):
    d = a - b
    d = 0.5 * (d + d.conj().T)
    return 0.5 * float(np.abs(np.linalg.eigvalsh(d)).sum())


def c5_problem(cavity_levels=25, n_ops=1):
    from sequencing_b200 import workloads
    return workloads.c5_transmon_two_cavities(cavity_levels=cavity_levels, n_ops=n_ops, sigma=3)


def test_c5_sparse_truth_is_converged_cpu():
    """The CPU truth on a small cousin of the system (3 x 5 x 5, N = 75): tighter tolerances move it by 3e-8, and the CPU
    oracle (QuTiP-4.x restatement, N^2 x N^2 Liouvillian - affordable at this size) agrees with it."""
    from helpers import oracle_solve
    bp = c5_problem(cavity_levels=5, n_ops=2)
    fin, ex, _ = sparse_truth(bp, rtol=1e-9, atol=1e-11)
    fin2, ex2, _ = sparse_truth(bp, rtol=1e-12, atol=1e-14)
    assert _trace_distance(fin, fin2) < 2e-7 and np.max(np.abs(ex - ex2)) < 2e-7
    ref = oracle_solve(bp, 0, atol=1e-12, rtol=1e-10, nsteps=100000)
    assert _trace_distance(fin2, np.asarray(ref.final_state).reshape(bp.N, bp.N)) < 1e-7
    assert np.max(np.abs(ex2 - np.real(np.array(ref.expect)))) < 1e-7


def _check_large(gpu_engine, bp, tag):
    # (the weighted RMS norm of both integrators averages over N^2 elements of which a few hundred carry the state: the
    # local error on those is ~ sqrt(N^2 / few hundred) above the requested tolerance - 1e-12 / 1e-13 it is)
    t0 = time.time()
    fin, ex, nrhs = sparse_truth(bp, rtol=1e-12, atol=1e-14)
    print(f"{tag}: sparse CPU truth, {nrhs} RHS evaluations in {time.time() - t0:.1f} s")
    out = gpu_engine.solve(bp)
    assert out.method == "large" and out.variant.startswith("large/diagonals") and out.status[0] == 0
    td, de = _trace_distance(out.final_state[0], fin), float(np.max(np.abs(out.expect[0] - ex)))
    bp.atol, bp.rtol, bp.nsteps = 1e-13, 1e-11, 100000
    tight = gpu_engine.solve(bp)
    assert tight.status[0] == 0
    td_t, de_t = _trace_distance(tight.final_state[0], fin), float(np.max(np.abs(tight.expect[0] - ex)))
    print(f"{tag} vs sparse CPU truth (DOP853 rtol 1e-12): default tolerances td {td:.2e} expect {de:.2e} | 1e-13/1e-11 td {td_t:.2e} expect {de_t:.2e}"
          f" ({int(tight.stats[0, 2])} RHS evaluations)")
    # matched tight tolerances: strict.  Default tolerances (atol 1e-8, rtol 1e-6) bound the LOCAL error per step; on these
    # strongly driven sigma = 3 pulses the accumulated error of either integrator is ~1e-5 (reported, loosely bounded)
    assert td_t <= 1e-7 and de_t <= 1e-7
    assert td <= 3e-5 and de <= 3e-5


@pytest.mark.gpu
def test_c5_n1875_large_path_against_sparse_cpu_truth(gpu_engine):
    bp = c5_problem()
    assert bp.N == 1875 and bp.B == 1 and bp.coeff.shape[-1] == 12
    _check_large(gpu_engine, bp, "C5 N=1875, first operation")


@pytest.mark.gpu
def test_c5_cousin_n300_four_operations_against_sparse_cpu_truth(gpu_engine):
    bp = c5_problem(cavity_levels=10, n_ops=4)
    assert bp.N == 300 and bp.coeff.shape[-1] == 48
    _check_large(gpu_engine, bp, "3 x 10 x 10 (N=300), four operations")
