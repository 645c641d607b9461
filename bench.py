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
NCU_SUMMARY = {"c2": "r01_diag_c2_cw_ncu_summary.json", "c4": "r01_diag_c4_final_ncu_summary.json",
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
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------
_ORACLE_BP = None


def _oracle_init(workload):
    global _ORACLE_BP
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    _ORACLE_BP, _ = build_workload(workload)


def _oracle_one(b):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_solve

    r = oracle_solve(_ORACLE_BP, int(b) % _ORACLE_BP.B)
    return float(np.real(r.expect[0][-1])) if r.expect else 0.0


class OraclePool:
    """The reference's CPU path (restated: oracle/qutip4_oracle.py) over all host cores,
    one trajectory per task (= qutip.parallel_map over sweep points)."""

    def __init__(self, workload, cores=None):
        import multiprocessing as mp

        self.cores = cores or os.cpu_count() or 1
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.cores, initializer=_oracle_init, initargs=(workload,))

    def run(self, n_traj, offset=0):
        t = time.perf_counter()
        self.pool.map(_oracle_one, range(offset, offset + n_traj), chunksize=1)
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bp, desc = build_workload(args.workload)
    pool = OraclePool(args.workload)
    sample = max(pool.cores * 2, 16)
    for w in range(args.warmup):
        pool.run(min(pool.cores, sample))
    times = [pool.run(sample, offset=k * sample) for k in range(args.steps)]
    pool.close()
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128 (f64)", "data": "synthetic",
        "config": {"workload": desc, **shape_of(bp), "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": pool.cores, "kind": "port",
                         "sample": f"{sample} trajectories/step of the same workload, one per task over {pool.cores} processes "
                                   "(QuTiP-4.x restatement: CSR Liouvillian SpMV + scipy ZVODE-Adams, atol 1e-8 rtol 1e-6 max_step dt)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sequencing_b200.solver import Engine

    from sequencing_b200.solver import _lib as seqlib
    bp, desc = build_workload(args.workload)   # every rank integrates its own shard of this size (weak scaling)
    if args.method != "auto":
        bp.method = {v: k for k, v in seqlib.METHOD_NAMES.items()}[args.method]
    rows_mode = args.workload == "c5"              # ONE system, rho row-sharded: strong scaling, collectives inside the solve
    if world > 1 and not rows_mode:
        # rank-dependent sweep values: scale the amplitudes a little so shards are distinct units
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
        from sequencing_b200.distributed import make_comm
        comm, keep = make_comm()

    def gather(outs):
        if world == 1 or rows_mode:
            return
        for key in ("expect", "final_state"):
            t = outs[key]
            buf = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(buf, t)

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
    n_units = bp.B if rows_mode else bp.B * world      # rows mode: the ranks share ONE trajectory
    value = n_units * args.steps / (total_ms * 1e-3)

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
                                   "(MEASURED_PEAKS.json has no FP64 entry): 36.99 TFLOP/s sustained",
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
            if sw:
                ex = sw["executed_flops_per_rhs"] * float(np.sum(stats[:, 2])) / (k_ms * 1e-3) / 1e12
                roof["structured"] = sw
                roof["executed_fp64_TFLOPs"] = ex
                roof["executed_fp64_frac_of_fma_peak"] = ex / FP64_FMA_PEAK_TFLOPS
                roof["note"] = ("`achieved` counts the ALGORITHMIC dense flops 8 N^3 (1 + 2 n_c) per RHS (SURVEY.md 8(d)); the kernel "
                                "AUTO picked exploits the operators' diagonal / row-sparse structure on the FP64 FMA pipe, so frac is "
                                "a speed-up over a dense tensor-core evaluation at peak, not a pipe utilisation - see `dense_kernel` "
                                "for the tensor-core kernel on the same inputs")
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
        cpu = None
        if world == 1 and not args.no_cpu:
            pool = OraclePool(args.workload)
            n = max(pool.cores * 4, 32)
            pool.run(pool.cores)
            t_cpu = pool.run(n)
            pool.close()
            cpu = {"value": n / t_cpu, "unit": UNIT, "cores": pool.cores, "kind": "port",
                   "sample": f"{n} of the {bp.B} trajectories, one per task over {pool.cores} processes "
                             "(oracle/qutip4_oracle.py: CSR Liouvillian + scipy ZVODE-Adams, reference options)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if rows_mode else "weak", "vs_baseline": None,
            "dtype": "c128 (f64)", "data": "synthetic",
            "config": {"workload": desc, **s, "B_per_gpu": bp.B, "atol": bp.atol, "rtol": bp.rtol, "max_step": bp.max_step,
                       "parallelism": (f"rho row-sharded x{world}, all-gather of the stage state per RHS" if rows_mode else f"batch-sharded x{world}"), "l2": "256 MiB L2 flush between timed steps; per-step scratch 300 MB > L2",
                       "method": out_h.method, "variant": out_h.variant.split(" [host solve")[0],
                       "e2e_host_pipeline": ("[host solve" + out_h.variant.split(" [host solve")[1]).strip("[]") if " [host solve" in out_h.variant else "single chunk",
                       "steps_accepted_mean": float(stats[:, 0].mean()), "steps_rejected_mean": float(stats[:, 1].mean())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
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
    ap.add_argument("--no-dense", action="store_true", help="skip the dense tensor-core kernel run reported in roofline.dense_kernel")
    ap.add_argument("--method", default="auto", help="force an engine kernel: auto|generic|dmma|dmma_stream|small|large|sparse|diag")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
