#!/usr/bin/env python
"""Benchmark of the hot path: batched Lindblad integration (``CompiledPulseSequence.run`` ->
``mesolve``) on the configuration BASELINE.json's metric is quoted on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

One "step" = one pass of the engine over one batch of synthetic sweep trajectories
(C2: transmon(3) (x) cavity(15), number-selective pi pulse, 32x32 amplitude x detuning sweep:
N=45, n_t=1000, n_ch=2, n_c=4, B=1024 per GPU).  With N GPUs every rank integrates its own
1024-trajectory shard (independent units -> weak scaling, no data-path collective) and the
expectation traces + final states are gathered with NCCL inside the timed step.

Prints ONE JSON line (rank 0).  ``value`` = trajectories/s with inputs resident in HBM;
``e2e`` = the same through the C-ABI call on HOST (pinned) buffers, H2D/D2H inside the timed
region; ``roofline`` = algorithmic FP64 flops of the integration kernel / its CUDA-event
duration against the measured DMMA peak; ``cpu_baseline`` = the CPU oracle (QuTiP-4.x
restatement: sparse Liouvillian + ZVODE) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "batched mesolve trajectories/sec at Hilbert dim N; % of FP64-DMMA/HBM roofline"
UNIT = "trajectories/s"
FP64_DMMA_PEAK_TFLOPS = 37.0  # measured on this pool's B200 by tools/fp64_peak.cu (profiles/r01_fp64_peak.json)
FP64_FMA_PEAK_TFLOPS = 33.2   # same tool, DFMA issue rate


def hbm_peak():
    """(GB/s, source): the driver-measured copy bandwidth, else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"


# the committed `ncu --set full` capture that belongs to each workload's integration kernel
NCU_SUMMARY = {"c2": "r02_diagblk_c2_ncu_summary.json", "c4": "r02_diagblk_c4_ncu_summary.json",
               "c1": "r01_small_c1_ncu_summary.json", "c3": "r01_small_c3_ncu_summary.json"}


def ncu_traffic(kernel, workload=None, rhs_evals=None):
    """dram bytes read+written per launch of `kernel` from the committed ncu --set full summary of this workload's
    kernel, or None.  The captures are shorter launches of the same kernel; when the summary records how many RHS
    evaluations the captured launch made, `bytes_per_launch_scaled` scales its DRAM bytes to this launch."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json")), reverse=True)
    pref = os.path.join(ROOT, "profiles", NCU_SUMMARY.get(workload or "", ""))
    if os.path.isfile(pref):
        paths = [pref] + [q for q in paths if q != pref]
    for path in paths:
        try:
            rows = json.load(open(path))
        except Exception:
            continue
        for r in rows:
            if kernel in r.get("kernel", ""):
                rd = [float(v) for k, v in r.items() if k.startswith("dram__bytes_read.sum")]
                wr = [float(v) for k, v in r.items() if k.startswith("dram__bytes_write.sum")]
                def scale_of(prefix):
                    unit = [k for k in r if k.startswith(prefix)]
                    return 1e6 if unit and "Mbyte" in unit[0] else (1e9 if unit and "Gbyte" in unit[0] else (1e3 if unit and "Kbyte" in unit[0] else 1.0))
                if rd and wr:
                    tot = rd[0] * scale_of("dram__bytes_read.sum") + wr[0] * scale_of("dram__bytes_write.sum")
                    out = {"bytes_per_launch": tot, "source": os.path.relpath(path, ROOT),
                           "note": "ncu capture of a shorter launch of the same kernel (see the file); not the bench launch"}
                    cap = r.get("capture") or {}
                    if rhs_evals and cap.get("rhs_total"):
                        out["bytes_per_launch_scaled"] = tot * float(rhs_evals) / float(cap["rhs_total"])
                        out["note"] += "; bytes_per_launch_scaled = captured DRAM bytes x (RHS evaluations of this launch / of the captured one)"
                    return out
    return None


def build_workload(name):
    from sequencing_b200 import workloads as w

    if name == "c2":
        bp = w.c2_selective(32, 32)
        desc = "C2 transmon(3)x cavity(15) selective pi pulse, 32x32 amp x detuning sweep"
    elif name == "c1":
        bp = w.c1_rabi(65536)
        desc = "C1 transmon(3) Rabi amplitude sweep, T1/T2 collapse ops"
    elif name == "c3":
        bp = w.c3_drag(4096)
        desc = "C3 transmon(4) DRAG sweep x 6 cardinal states"
    elif name == "c4":
        bp = w.c4_two_transmons_bus(512)
        desc = "C4 two transmons(3) + bus cavity(10), full collapse-operator set, batch 512"
    elif name == "c5":
        n_ops = int(os.environ.get("SEQ_C5_OPS", "5"))
        bp = w.c5_transmon_two_cavities(n_ops=n_ops)
        desc = (f"C5 transmon(3) x cavity(25) x cavity(25), N=1875, ONE system, first {n_ops} of the sequence's operations "
                f"({bp.coeff.shape[-1]} samples); rho row-sharded over the GPUs (strong scaling)")
    elif name == "c2small":
        bp = w.c2_selective(4, 4, sigma=25)
        desc = "C2-small (smoke) transmon(3)x cavity(15), 4x4 sweep, n_t=100"
    else:
        raise SystemExit(f"unknown workload {name}")
    return bp, desc


def shape_of(bp):
    return dict(N=bp.N, B=bp.B, n_t=int(bp.coeff.shape[-1]), n_ch=int(bp.Hk.shape[0]),
                n_c=int(bp.Lc.shape[0] + bp.Ltd.shape[0]), n_e=int(bp.E.shape[0]))


def algorithmic(bp, stats):
    """F_rhs = 8 N^3 (1 + 2 n_c) per RHS evaluation; bytes per trajectory =
    8 n_ch n_t + 16 N^2 + 8 n_e n_t + 16 N^2."""
    s = shape_of(bp)
    f_rhs = 8.0 * s["N"] ** 3 * (1 + 2 * s["n_c"])
    rhs_evals = float(np.sum(stats[:, 2]))
    bytes_traj = 8.0 * s["n_ch"] * s["n_t"] + 32.0 * s["N"] ** 2 + 8.0 * s["n_e"] * s["n_t"]
    return f_rhs * rhs_evals, bytes_traj * s["B"], f_rhs


def tile_sparsity(bp, method):
    """Fraction of the dense 8x4 operator fragments the tensor-core kernels actually multiply (same rule
    as dmma_masks_kernel: a fragment is skipped iff every entry is exactly zero), and the real DMMA work
    it implies.  Reported next to the algorithmic figure so that frac > 1 is not misread."""
    N = bp.N
    if method == "dmma":
        NP = 8 * ((N + 7) // 8)
    elif method == "dmma_stream":
        NP = 16 * (((N + 7) // 8 + 1) // 2)
    else:
        return None

    def frags(M):
        P = np.zeros((NP, NP), complex)
        P[:N, :N] = M
        nz = np.abs(P).reshape(NP // 8, 8, NP // 4, 4).max(axis=(1, 3)) > 0
        return int(nz.sum())

    dense = (NP // 8) * (NP // 4)
    Ls = list(bp.Lc) + list(bp.Ltd if bp.Ltd is not None else [])
    G = np.abs(bp.H0 if bp.H0.ndim == 2 else bp.H0[0]) + sum(np.abs(h) for h in bp.Hk) + sum(np.abs(L.conj().T @ L) for L in Ls)
    g, ls = frags(G), [frags(L) for L in Ls]
    executed = g + 2 * sum(ls)
    full = dense * (1 + 2 * len(Ls))
    # executed real flops per algorithmic flop: padding (NP/N)^3, Karatsuba 3/4, fragment density
    return {"fragments_dense_per_operator": dense, "G_fragments": g, "L_fragments": ls,
            "dmma_fragment_density": executed / full,
            "executed_over_algorithmic_flops": (NP / N) ** 3 * 0.75 * executed / full,
            "note": "operators are Kronecker-embedded ladder/number operators: all-zero 8x4 fragments are skipped (exact); "
                    "complex products use 3 real DMMA chains (Karatsuba) instead of 4"}


def structured_work(bp, method):
    """What the structure-exploiting kernels (diag / sparse) really execute per RHS: complex multiply-adds per
    owned element = 2 (diagonals | row width of G) + sum_l (diagonals | row width of L_l)^2, over the N (N+1) / 2
    elements of the upper triangle, 8 real flops each.  Reported next to the algorithmic (dense) figure so that
    frac > 1 is not misread as tensor-pipe utilisation."""
    if method not in ("diag", "sparse"):
        return None
    N = bp.N
    Ls = list(bp.Lc if bp.Lc is not None else []) + list(bp.Ltd if bp.Ltd is not None else [])
    G = np.abs(bp.H0 if bp.H0.ndim == 2 else bp.H0[0]) + sum(np.abs(h) for h in bp.Hk) + sum(np.abs(L.conj().T @ L) for L in Ls)

    def width(M):
        if method == "diag":
            i, j = np.nonzero(M)
            return len(set((j - i).tolist()))
        return int((M != 0).sum(axis=1).max())

    wg, wl = width(G), [width(np.abs(L)) for L in Ls]
    macs = 2 * wg + sum(w * w for w in wl)
    flops = 8.0 * macs * N * (N + 1) / 2
    return {"G_terms": wg, "L_terms": wl, "complex_macs_per_element": macs, "executed_flops_per_rhs": flops,
            "note": ("operators stored by diagonals" if method == "diag" else "operators stored as ELL rows")
                    + "; rho Hermitian -> upper triangle only; FP64 FMA pipe, no tensor cores"}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU legs: the reference's path (qutip.mesolve when a QuTiP 4.x is importable on the box, else the
# restated oracle) timed serially (the calibration.py:120-125 loop, one core) and over all host cores
# (one trajectory per task = qutip.parallel_map over sweep points).  Used by --impl reference and, in a
# child process started BEFORE this process touches CUDA, by the cpu_baseline object of our arm - one
# code path, so the two numbers are the same measurement.
# ------------------------------------------------------------------------------------------
_ORACLE_BP = None
_USE_QUTIP = False


def qutip_probe():
    """A real QuTiP 4.x (prefer it when the box has it - e.g. under baseline/_ref)?"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.append(ref)
    try:
        import qutip  # noqa: F401
        return str(qutip.__version__).startswith("4.")
    except Exception:
        return False


def _oracle_init(workload, use_qutip=False):
    global _ORACLE_BP, _USE_QUTIP
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    try:      # the BLAS pool was sized when numpy was imported: cap it after the fact as well
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    _ORACLE_BP, _ = build_workload(workload)
    _USE_QUTIP = bool(use_qutip) and qutip_probe()


def _qutip_one(bp, b):
    """The reference's own call (basic.py:476-512) on trajectory b: qutip.mesolve, list-format H with array
    coefficients, Options(max_step=dt), e_ops - QuTiP 4.x only."""
    import qutip
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_inputs
    H, s0, tlist, c_ops, e_ops, grid = oracle_inputs(bp, b)
    Hq = [qutip.Qobj(H[0])] + [[qutip.Qobj(op), np.asarray(c, float)] for op, c in H[1:]]
    cq = [qutip.Qobj(L) if not isinstance(L, tuple) else [qutip.Qobj(L[0]), np.asarray(L[1], float)] for L in c_ops]
    opt = qutip.Options(atol=bp.atol, rtol=bp.rtol, nsteps=max(int(bp.nsteps), 1000), max_step=bp.max_step, store_states=False)
    r = qutip.mesolve(Hq, qutip.Qobj(s0), grid, cq, [qutip.Qobj(E) for E in e_ops], options=opt)
    return float(np.real(r.expect[0][-1])) if r.expect else 0.0


def _oracle_one(b):
    b = int(b) % _ORACLE_BP.B
    if _USE_QUTIP:
        return _qutip_one(_ORACLE_BP, b)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_solve

    r = oracle_solve(_ORACLE_BP, b)
    return float(np.real(r.expect[0][-1])) if r.expect else 0.0


class OraclePool:
    """All host cores, one trajectory per task (= qutip.parallel_map over sweep points)."""

    def __init__(self, workload, cores=None, use_qutip=False):
        import multiprocessing as mp

        self.cores = cores or os.cpu_count() or 1
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_oracle_init, initargs=(workload, use_qutip))

    def run(self, n_traj, offset=0):
        t = time.perf_counter()
        self.pool.map(_oracle_one, range(offset, offset + n_traj), chunksize=1)
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def spread_indices(B, n):
    """n trajectory indices spread over the whole sweep (every region of the parameter grid, not its first rows)."""
    n = min(n, B)
    return [int(x) for x in np.unique(np.round(np.linspace(0, B - 1, n)).astype(int))]


def cpu_legs(workload, steps=1, warmup=1, per_core=4):
    """Serial and parallel timing of the reference's CPU path on a bounded sample spread over the sweep."""
    bp, desc = build_workload(workload)
    use_qutip = qutip_probe()
    kind = "qutip" if use_qutip else "port"
    what = ("qutip.mesolve (QuTiP 4.x found on this box)" if use_qutip else
            "QuTiP-4.x restatement oracle/qutip4_oracle.py: CSR Liouvillian SpMV + scipy ZVODE-Adams") + ", atol 1e-8 rtol 1e-6 max_step dt"
    # serial: the calibration.py:120-125 loop on one core, a few sweep points
    _oracle_init(workload, use_qutip)
    idx = spread_indices(bp.B, 6)
    _oracle_one(idx[0])
    t = time.perf_counter()
    for b in idx:
        _oracle_one(b)
    t_serial = time.perf_counter() - t
    serial = {"value": len(idx) / t_serial, "unit": UNIT, "cores": 1, "sample": f"{len(idx)} sweep points spread over the {bp.B}, one after the other in one process"}
    pool = OraclePool(workload, use_qutip=use_qutip)
    n = max(pool.cores * per_core, 32)
    stride = max(1, bp.B // n)
    for w in range(warmup):
        pool.run(pool.cores)
    times = []
    for k in range(steps):
        t = time.perf_counter()
        pool.pool.map(_oracle_one, [(k + q * stride) % bp.B for q in range(n)], chunksize=1)
        times.append(time.perf_counter() - t)
    pool.close()
    par = {"value": n * steps / sum(times), "unit": UNIT, "cores": pool.cores,
           "sample": f"{n} sweep points per step spread over the {bp.B} (stride {stride}), one per task over {pool.cores} processes"}
    return {"value": par["value"], "unit": UNIT, "cores": pool.cores, "kind": kind, "sample": par["sample"] + "; " + what,
            "serial": serial, "parallel": par, "ms_per_step": 1e3 * sum(times) / steps, "sample_per_step": n}, bp, desc


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu, bp, desc = cpu_legs(args.workload, steps=args.steps, warmup=args.warmup, per_core=2)
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cpu.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": {"workload": desc, **shape_of(bp), "sample_per_step": cpu.pop("sample_per_step")},
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_child(workload):
    """cpu_baseline of our arm: the same legs in a child process, started before this process initialises CUDA."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1"],
                             capture_output=True, text=True, timeout=900, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as exc:      # the baseline is a reported number, never a reason to lose the bench line
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"cpu legs failed: {exc!r}"}


# strict matched tolerance per workload: the loosest pair at which the oracle (ZVODE) is itself converged to < 1e-7
STRICT_TOL = {"c1": (1e-10, 1e-8), "c3": (1e-10, 1e-8), "c4": (1e-10, 1e-8), "c2": (1e-12, 1e-10), "c2small": (1e-12, 1e-10)}


def parity_block(eng, bp, workload, n_samples, method=None):
    """Engine vs the CPU oracle at MATCHED atol/rtol on trajectories spread over the sweep: pass fraction of the
    outputs (final-state trace distance and expectation-trace max-abs error <= 1e-6) at the reference's default
    tolerances and at the strict pair, both attributed against the converged (strict) oracle run.  The oracle is the
    checker here, nothing it computes is timed or shipped.  The 64-trajectory records of tests/test_parity_sampled.py
    are under profiles/r02_parity_*.json."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_many, parity_fractions, take_one
    import copy
    default, strict = (bp.atol, bp.rtol), STRICT_TOL.get(workload, (1e-10, 1e-8))
    idx = spread_indices(bp.B, n_samples)
    sub = copy.copy(bp)
    for name in ("coeff", "rate", "state0", "coeff_scale"):
        arr = getattr(sub, name)
        if arr is not None and arr.shape[0] == bp.B:
            setattr(sub, name, np.ascontiguousarray(arr[idx]))
    if method is not None:
        sub.method = method      # (the kernel the bench timed: AUTO would keep the plain diagonal kernel for so small a batch)
    outs = {}
    for atol, rtol in (default, strict):
        sub.atol, sub.rtol = atol, rtol
        outs[(atol, rtol)] = eng.solve(sub)
    refs = oracle_many(bp, idx, [default, strict])
    rep = parity_fractions(sub, list(range(len(idx))), outs, refs, converged=strict)
    d, st = rep["atol=%g,rtol=%g" % default], rep["atol=%g,rtol=%g" % strict]
    return {"tolerance": 1e-6, "sampled_trajectories": len(idx), "outputs_per_trajectory": 1 + int(bp.E.shape[0]),
            "default": {"atol": default[0], "rtol": default[1], "pass_fraction_engine_vs_oracle": d["fraction"],
                        "worst_trace_distance": d["worst_trace_distance"], "worst_expect_abs_err": d["worst_expect_abs_err"],
                        "pass_fraction_engine_vs_converged": d["engine_vs_converged"]["fraction"],
                        "pass_fraction_oracle_vs_converged": d["oracle_vs_converged"]["fraction"]},
            "strict": {"atol": strict[0], "rtol": strict[1], "pass_fraction_engine_vs_oracle": st["fraction"],
                       "worst_trace_distance": st["worst_trace_distance"], "worst_expect_abs_err": st["worst_expect_abs_err"]},
            "oracle": "oracle/qutip4_oracle.py (QuTiP-4.x restatement, parity unpinned against QuTiP itself - see DESIGN.md 2)",
            "records": "profiles/r02_parity_{c1,c2,c3,c4}.json: 64 trajectories per config (tests/test_parity_sampled.py)"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    if world == 1 and not args.no_cpu and args.workload != "c5":      # (the oracle's N^2 x N^2 Liouvillian does not exist at N = 1875: tests/test_c5_large.py
        cpu = cpu_baseline_child(args.workload)                         #  holds the sparse-operator CPU truth for C5)      # before torch / CUDA come up in this process
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sequencing_b200.solver import Engine

    from sequencing_b200.solver import _lib as seqlib
    bp, desc = build_workload(args.workload)   # every rank integrates its own shard of this size (weak scaling)
    if args.method != "auto":
        bp.method = {v: k for k, v in seqlib.METHOD_NAMES.items()}[args.method]
    rows_mode = args.workload == "c5"              # ONE system, rho row-sharded: strong scaling, collectives inside the solve
    strong = args.scaling == "strong" and not rows_mode
    B_total = bp.B
    if world > 1 and strong:
        # the NAMED sweep (B trajectories in total) cut into contiguous slices, one per GPU
        lo, hi = rank * bp.B // world, (rank + 1) * bp.B // world
        import copy as _cp
        bp = _cp.copy(bp)
        for name in ("coeff", "rate", "state0", "coeff_scale"):
            arr = getattr(bp, name)
            if arr is not None and arr.shape[0] == B_total:
                setattr(bp, name, np.ascontiguousarray(arr[lo:hi]))
    elif world > 1 and not rows_mode:
        # weak scaling: every rank integrates a full-size sweep of its own (amplitudes scaled a little per rank so that the
        # shards are distinct units)
        bp.coeff = bp.coeff * (1.0 + 1e-3 * rank)
    eng = Engine.get(local)
    dev = torch.device("cuda", local)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    comm = keep = None
    if rows_mode and world > 1:
        # NCCL issued by the engine itself (seq_engine_nccl_init); --hooks keeps the torch.distributed callbacks of round 1
        if args.hooks:
            from sequencing_b200.distributed import make_comm
            comm, keep = make_comm(stream=eng.stream)
        else:
            eng.init_native_comm()

    gathered = {}

    def gather(outs):
        """ONE NCCL all-gather per step: expectation traces and final states packed into one buffer (equal slices)."""
        if world == 1 or rows_mode:
            return
        ex, fs = outs["expect"].reshape(-1), torch.view_as_real(outs["final_state"]).reshape(-1)
        n_loc = torch.tensor([ex.numel() + fs.numel()], device=dev)
        if strong:      # slices may differ by one trajectory: pad to the largest
            dist.all_reduce(n_loc, op=dist.ReduceOp.MAX)
        n = int(n_loc.item()) if strong else ex.numel() + fs.numel()
        send = gathered.get("send")
        if send is None or send.numel() != n:
            send = gathered["send"] = torch.zeros(n, dtype=torch.float64, device=dev)
            gathered["recv"] = torch.empty(n * world, dtype=torch.float64, device=dev)
        send[:ex.numel()] = ex
        send[ex.numel():ex.numel() + fs.numel()] = fs
        dist.all_gather_into_tensor(gathered["recv"], send)

    # ---- device-resident value -----------------------------------------------------------
    tens = eng.to_device(bp)
    torch.cuda.synchronize()
    kernel_ms, step_ms, launches = [], [], 0
    def solve_dev():
        if rows_mode:
            return eng.solve_rows_sharded(bp, comm, tens=tens, to_host=False)
        return eng.solve_device(tens, bp)

    for w in range(args.warmup):
        res = solve_dev()
        gather(res["out"])
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush_buf.fill_(k)            # L2 flush between timed iterations (outside the events)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = solve_dev()
        gather(res["out"])
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kernel_ms.append(eng.kernel_ms())
        launches += eng.last_launches
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    stats = res["out"]["stats"].cpu().numpy()
    status = res["out"]["status"].cpu().numpy()
    assert np.all(status == 0), "integration failed for some trajectories"
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    n_units = bp.B if rows_mode else (B_total if strong else bp.B * world)      # rows mode: the ranks share ONE trajectory
    value = n_units * args.steps / (total_ms * 1e-3)

    # ---- multi-GPU correctness: what arrived from rank 1 equals a single-GPU solve of rank 1's shard on rank 0 -------
    mgpu_check = None
    if world > 1 and not rows_mode and rank == 0:
        import copy as _cp
        other = _cp.copy(build_workload(args.workload)[0])
        if args.method != "auto":
            other.method = bp.method
        if strong:
            lo1, hi1 = 1 * B_total // world, 2 * B_total // world
            for name in ("coeff", "rate", "state0", "coeff_scale"):
                arr = getattr(other, name)
                if arr is not None and arr.shape[0] == B_total:
                    setattr(other, name, np.ascontiguousarray(arr[lo1:hi1]))
        else:
            other.coeff = other.coeff * (1.0 + 1e-3 * 1)
        chk = eng.solve_device(eng.to_device(other), other)["out"]
        torch.cuda.synchronize()
        n_ex, n_fs = chk["expect"].numel(), 2 * chk["final_state"].numel()
        seg = gathered["recv"][gathered["send"].numel():2 * gathered["send"].numel()]
        d_ex = float((seg[:n_ex] - chk["expect"].reshape(-1)).abs().max().item())
        d_fs = float((seg[n_ex:n_ex + n_fs] - torch.view_as_real(chk["final_state"]).reshape(-1)).abs().max().item())
        mgpu_check = {"what": "rank 1's slice of the NCCL-gathered outputs vs a single-GPU solve of the same shard on rank 0",
                      "max_abs_diff_expect": d_ex, "max_abs_diff_final_state": d_fs, "ok": bool(d_ex < 1e-12 and d_fs < 1e-12)}
        assert mgpu_check["ok"], mgpu_check

    # ---- end-to-end through the C ABI on pinned host buffers ---------------------------------
    host = eng._host_arrays(bp)
    pinned = {}
    for name, arr in host.items():
        t = torch.from_numpy(arr).pin_memory() if arr.size else torch.from_numpy(arr)
        pinned[name] = t
    import copy
    bp_h = copy.copy(bp)
    for name in ("H0", "Hk", "coeff", "Lc", "Ltd", "rate", "state0", "E", "out_idx"):
        if name in pinned:
            setattr(bp_h, name, pinned[name].numpy())
    h2d = sum(int(v.numel() * v.element_size()) for v in pinned.values())
    def solve_host():
        if rows_mode:                      # public API of the sharded mode: host arrays in, host arrays out
            return eng.solve_rows_sharded(bp_h, comm)
        return eng.solve_host_abi(bp_h, pinned_out=True)    # returns after D2H (into pinned buffers) + stream sync

    solve_host()  # warm the staging buffers
    barrier()
    e2e_t = []
    for k in range(max(1, min(args.steps, 3))):
        flush_buf.fill_(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out_h = solve_host()
        e2e_t.append(time.perf_counter() - t0)
        launches += eng.last_launches
    d2h = int(out_h.expect.nbytes + out_h.final_state.nbytes + out_h.status.nbytes + out_h.stats.nbytes)
    t_e2e = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_units * len(e2e_t) / float(t_e2e.item())

    if rank == 0:
        flops, bytes_alg, f_rhs = algorithmic(bp, stats)
        k_ms = float(np.mean(kernel_ms))
        achieved = flops / (k_ms * 1e-3) / 1e12
        s = shape_of(bp)
        structured_large = out_h.method == "large" and out_h.variant.startswith("large/diagonals")
        kname = ("large_diag_rhs_kernel" if structured_large else "zgemm_dmma_kernel") if out_h.method == "large" else f"solve_{out_h.method}_kernel"
        if out_h.variant.startswith("diagblk"):
            kname = "solve_diagblk_kernel"
        herm_large = structured_large and "Hermitian tiles" in out_h.variant
        if herm_large:
            kname = "lhq_stage_kernel" if "outer factor of dimension" in out_h.variant else "lh_stage_kernel"
        peak_scale = world if rows_mode else 1     # rows mode: the algorithmic flops of the ONE system are spread over all GPUs
        if out_h.method == "small":
            # N <= 6: below one DMMA tile -> the HBM roofline, FP64-FMA ceiling noted
            peak, src = hbm_peak()
            gbps = bytes_alg / (k_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": gbps, "peak": peak, "unit": "GB/s", "frac": gbps / peak, "traffic": ncu_traffic(kname, args.workload),
                    "kernel": kname, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (total_ms / args.steps),
                    "bytes_per_trajectory": bytes_alg / s["B"], "trajectories_per_launch": s["B"], "peak_source": src,
                    "fp64_fma_algorithmic_TFLOPs": achieved, "fp64_fma_peak_TFLOPs": FP64_FMA_PEAK_TFLOPS,
                    "note": "FP64-FMA-issue bound, not HBM bound: the kernel executes (1+n_g) N^4 real FMAs per RHS "
                            "(real superoperator form) instead of the 8 N^3 (1+2 n_c) algorithmic flops"}
        else:
            roof = {"bound": "tensor", "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS * peak_scale, "unit": "TFLOP/s",
                    "frac": achieved / (FP64_DMMA_PEAK_TFLOPS * peak_scale),
                    "traffic": ncu_traffic(kname, args.workload, float(np.sum(stats[:, 2]))),
                    "kernel": kname, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (total_ms / args.steps),
                    "flops_per_rhs": f_rhs, "rhs_evals_per_launch": float(np.sum(stats[:, 2])),
                    "peak_source": "FP64 DMMA.8x8x4 issue-rate microbenchmark tools/fp64_peak.cu on this pool's B200 "
                                   "(MEASURED_PEAKS.json has no FP64 entry): 36.99 TFLOP/s sustained; FP64 FMA 33.2 TFLOP/s (same tool)",
                    "hbm_algorithmic_GBps": bytes_alg / (k_ms * 1e-3) / 1e9}
            sp = tile_sparsity(bp, out_h.method)
            if sp:
                roof["tile_sparsity"] = sp
                roof["executed_dmma_TFLOPs"] = achieved * sp["executed_over_algorithmic_flops"]
                roof["executed_dmma_frac_of_peak"] = roof["executed_dmma_TFLOPs"] / FP64_DMMA_PEAK_TFLOPS
            sw = structured_work(bp, "diag" if structured_large else out_h.method)
            if structured_large:
                # the structured RHS of the large path is HBM-bound: 16 N^2 bytes of state read + 16 N^2 written per RHS
                peak_h, src_h = hbm_peak()
                gb = 32.0 * s["N"] ** 2 * float(np.sum(stats[:, 2])) / (k_ms * 1e-3) / 1e9
                roof["hbm"] = {"bound": "hbm", "achieved": gb, "peak": peak_h * peak_scale, "unit": "GB/s", "frac": gb / (peak_h * peak_scale),
                               "bytes_per_rhs": 32.0 * s["N"] ** 2, "peak_source": src_h,
                               "note": "whole stepper time (RHS + stage combinations + error norm + outputs) against the RHS kernel's compulsory bytes"}
                if herm_large:
                    # what the Hermitian-tile stepper moves by construction (kernel_large_herm.cuh): h = one packed half vector;
                    # a 4(3) step streams 37 h (stage input 4 h, fused stages 7 + 8 + 10 + 8 h), a 5(4) step 59 h
                    nt = (s["N"] + 31) // 32
                    h_b = 16.0 * 1024 * nt * (nt + 1) / 2
                    if "outer factor of dimension" in out_h.variant:      # 16 x 16 tiles of the x-space, Q x Q elements per pair (kernel_large_hermq.cuh)
                        q_loc = int(out_h.variant.split("outer factor of dimension")[1].split()[0])
                        nxt = (s["N"] // q_loc + 15) // 16
                        h_b = 16.0 * 256 * q_loc * q_loc * nxt * (nxt + 1) / 2
                    rhs_all, att_all = float(np.sum(stats[:, 2])), float(np.sum(stats[:, 0]) + np.sum(stats[:, 1]))
                    f5 = min(1.0, max(0.0, (rhs_all / max(att_all, 1.0) - 4.0) / 2.0))
                    per_rhs = h_b * ((1.0 - f5) * 37.0 / 4.0 + f5 * 59.0 / 6.0)
                    roof["hbm"].update({"executed_bytes_per_rhs_model": per_rhs, "executed_GBps_model": per_rhs * rhs_all / (k_ms * 1e-3) / 1e9,
                                        "executed_frac_model": per_rhs * rhs_all / (k_ms * 1e-3) / 1e9 / (peak_h * peak_scale),
                                        "model": "packed half vectors of 16 B x 1024 x tiles(ti<=tj): stage input 4 h, fused stage kernels 7/8/10/8 h (4(3) pair) or "
                                                 "7/8/9/10/12/9 h (5(4) pair); ncu dram__bytes of the four stage kernels of a 4(3) step (32 x 32 tile variant) agrees (profiles/r02_c5_herm_stage_ncu_summary.json)"})
            if sw:
                # The kernel that ran is a structure-exploiting FP64-FMA kernel: `frac` is what it EXECUTES against the
                # peak of the pipe it uses; the dense-equivalent figure (SURVEY.md 8(d) algorithmic flops against the DMMA
                # peak) is a speed-up, not a utilisation, and is reported as such.
                fma_rhs = int(eng.lib.seq_engine_last_fma_per_rhs(eng.handle))
                if fma_rhs > 0:
                    sw = dict(sw, executed_flops_per_rhs=2.0 * fma_rhs,
                              executed_count="FP64 multiply-adds of the right-hand-side terms counted by the kernel's host plan from its term "
                                             "lists (seq_engine_last_fma_per_rhs; terms a thread skips statically are not counted; stage "
                                             "combinations, error norm and outputs are not counted either)")
                ex = sw["executed_flops_per_rhs"] * float(np.sum(stats[:, 2])) / (k_ms * 1e-3) / 1e12
                extra = int(eng.lib.seq_engine_last_step_extra_ops(eng.handle))
                if extra > 0:
                    # everything the FP64 pipe executes: + stage combinations and error estimate of every step (counted per 4(3)
                    # step; the few 5(4) steps execute more).  ncu's sm__pipe_fp64_cycles_active of the same launch is in
                    # profiles/r02_diagblk_c2_ncu_summary.json (17.2 % before the tail split / owned-half changes).
                    steps_all = float(np.sum(stats[:, 0]) + np.sum(stats[:, 1]))
                    roof["fp64_pipe_instructions_TFLOP_equiv"] = ex + 2.0 * extra * steps_all / (k_ms * 1e-3) / 1e12
                    roof["fp64_pipe_frac_counted"] = roof["fp64_pipe_instructions_TFLOP_equiv"] / FP64_FMA_PEAK_TFLOPS
                ex_all = roof.get("fp64_pipe_instructions_TFLOP_equiv", ex)
                roof.update({"bound": "fp64-fma", "achieved": ex_all, "peak": FP64_FMA_PEAK_TFLOPS * peak_scale, "frac": ex_all / (FP64_FMA_PEAK_TFLOPS * peak_scale),
                             "frac_rhs_terms_only": ex / (FP64_FMA_PEAK_TFLOPS * peak_scale),
                             "structured": sw, "executed_fp64_TFLOPs": ex, "executed_fp64_frac_of_fma_peak": ex / FP64_FMA_PEAK_TFLOPS,
                             "dense_equivalent_TFLOPs": achieved, "dense_equivalent_speedup": achieved / (FP64_DMMA_PEAK_TFLOPS * peak_scale),
                             "note": "bound = the FP64 FMA pipe (no tensor cores: the reference's operators are sums of a few diagonals and the "
                                     "kernel AUTO picks multiplies only those); frac = FP64 pipe instructions the kernel executes (right-hand-side "
                                     "terms + stage combinations + error estimate, counted by its host plan, x 2 flop) / time / FMA peak - ncu's "
                                     "sm__pipe_fp64_cycles_active of the same launch agrees (profiles/r02_diagblk_c2_ncu_summary.json); "
                                     "frac_rhs_terms_only counts the multiply-adds of the right-hand-side terms alone.  dense_equivalent_speedup "
                                     "= algorithmic dense flops 8 N^3 (1 + 2 n_c) per RHS (SURVEY.md 8(d)) / time / DMMA peak: how much faster "
                                     "than a dense tensor-core evaluation at 100 % of its peak - not a utilisation.  `dense_kernel` is the "
                                     "tensor-core kernel timed on the same inputs.  ncu pipe utilisation of this kernel: profiles/"})
                if structured_large:
                    # the large path's structured stepper is HBM-bound (a few FP64 multiply-adds per 16-byte element): the top-level
                    # roofline is the HBM one - SURVEY.md 8(d)'s 32 N^2 bytes per RHS against the WHOLE stepper's time; the FP64 figures stay as extras
                    roof.update({"bound": "hbm", "achieved": roof["hbm"]["achieved"], "peak": roof["hbm"]["peak"], "unit": "GB/s", "frac": roof["hbm"]["frac"],
                                 "peak_source": roof["hbm"]["peak_source"]})
                if world == 1 and not args.no_dense and not rows_mode:
                    # the dense tensor-core formulation on the same workload (a quarter of the batch), for the DMMA roofline
                    import copy as _copy
                    nb = max(1, bp.B // 4)
                    bpd = _copy.copy(bp)
                    for name in ("coeff", "rate", "state0", "coeff_scale"):
                        a = getattr(bpd, name)
                        if a is not None and a.shape[0] == bp.B:
                            setattr(bpd, name, a[:nb])
                    bpd.method = seqlib.METHOD_DMMA if bp.N <= 48 else (seqlib.METHOD_DMMA_STREAM if bp.N <= 96 else seqlib.METHOD_LARGE)
                    tens_d = eng.to_device(bpd)
                    eng.solve_device(tens_d, bpd); torch.cuda.synchronize()
                    rd = eng.solve_device(tens_d, bpd); torch.cuda.synchronize()
                    launches += 2 * eng.last_launches
                    kd = eng.kernel_ms()
                    std = rd["out"]["stats"].cpu().numpy()
                    fd = f_rhs * float(np.sum(std[:, 2])) / (kd * 1e-3) / 1e12
                    dname = seqlib.METHOD_NAMES[bpd.method]
                    dsp = tile_sparsity(bpd, dname)
                    roof["dense_kernel"] = {"method": dname, "trajectories": nb, "kernel_ms": kd, "trajectories_per_s": nb / (kd * 1e-3),
                                            "algorithmic_TFLOPs": fd, "frac_of_dmma_peak": fd / FP64_DMMA_PEAK_TFLOPS,
                                            "executed_dmma_frac_of_peak": (fd * dsp["executed_over_algorithmic_flops"] / FP64_DMMA_PEAK_TFLOPS) if dsp else None}
        parity = None
        if world == 1 and args.parity_samples > 0 and not rows_mode and args.workload in STRICT_TOL:
            try:
                parity = parity_block(eng, bp, args.workload, args.parity_samples, seqlib.METHOD_DIAG if out_h.method == "diag" else None)
                parity["engine_kernel"] = kname
            except Exception as exc:
                parity = {"error": repr(exc)}
        if world == 1 and herm_large and parity is None:
            # C5: the oracle's N^2 x N^2 Liouvillian does not exist at N = 1875 (tests/test_c5_large.py holds the sparse-operator CPU
            # truth); what a bench run can check is the benched kernels against the full-matrix stepper of the same path on the
            # sequence's first operation - same steps, same numbers
            try:
                import sequencing_b200.workloads as _w
                outs = {}
                for mode in ("herm", "full"):
                    one = _w.c5_transmon_two_cavities(n_ops=1)
                    one.method = seqlib.METHOD_LARGE
                    if mode == "full":
                        os.environ["SEQ_LARGE_FULL"] = "1"
                    try:
                        outs[mode] = eng.solve(one)
                    finally:
                        os.environ.pop("SEQ_LARGE_FULL", None)
                parity = {"what": "Hermitian-tile stepper vs the full-matrix stepper of the large path, first operation of the C5 sequence (GPU vs GPU; "
                                  "the CPU truth at N = 1875 is tests/test_c5_large.py)",
                          "variants": [outs["herm"].variant, outs["full"].variant],
                          "max_abs_final_state": float(np.max(np.abs(outs["herm"].final_state - outs["full"].final_state))),
                          "max_abs_expect": float(np.max(np.abs(outs["herm"].expect - outs["full"].expect))),
                          "same_steps": bool(np.array_equal(outs["herm"].stats[:, :3], outs["full"].stats[:, :3])),
                          "engine_kernel": kname}
            except Exception as exc:
                parity = {"error": repr(exc)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if (rows_mode or strong) else "weak", "vs_baseline": None,
            "dtype": "c128 (f64)", "data": "synthetic",
            "config": {"workload": desc, **s, "B": B_total if strong else s["B"], "B_per_gpu": bp.B, "atol": bp.atol, "rtol": bp.rtol, "max_step": bp.max_step,
                       "parallelism": (f"rho row-sharded x{world}, all-gather of the stage state per RHS" if rows_mode else f"batch-sharded x{world}, one packed NCCL all-gather of expect + final states per step"), "l2": "256 MiB L2 flush between timed steps; per-step scratch 300 MB > L2",
                       "method": out_h.method, "variant": out_h.variant.split(" [host solve")[0],
                       "e2e_host_pipeline": ("[host solve" + out_h.variant.split(" [host solve")[1]).strip("[]") if " [host solve" in out_h.variant else "single chunk",
                       "steps_accepted_mean": float(stats[:, 0].mean()), "steps_rejected_mean": float(stats[:, 1].mean())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "mgpu_check": mgpu_check,
            "clocks": clocks,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--hooks", action="store_true", help="c5 on N > 1 GPUs: collectives through torch.distributed callbacks instead of the engine's own NCCL communicator")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: every GPU its own full sweep (weak) or the named sweep cut into N slices (strong)")
    ap.add_argument("--parity-samples", type=int, default=8, help="trajectories of the sweep checked against the CPU oracle for the `parity` block (0: skip)")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense tensor-core kernel run reported in roofline.dense_kernel")
    ap.add_argument("--method", default="auto", help="force an engine kernel: auto|generic|dmma|dmma_stream|small|large|sparse|diag")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
