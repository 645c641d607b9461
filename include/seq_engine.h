/*
 * seq_engine.h - C ABI of the B200-native batched Lindblad / Schroedinger integrator.
 *
 * This is the drop-in boundary for ONE hot path of sequencing-dev/sequencing: the time
 * evolution that `CompiledPulseSequence.run()` hands to a third-party solver at
 *     sequencing/sequencing/basic.py:504-512   qutip.mesolve(H, init_state, tlist, c_ops, e_ops, options)
 *     sequencing/sequencing/basic.py:481-485   qutip.Options(): max_step = dt, store_states = True
 *     sequencing/sequencing/basic.py:500       qutip.QobjEvo(H, tlist=times)   (only_final_state)
 *     sequencing/sequencing/basic.py:560-567   qutip.propagator(H, times, c_op_list=c_ops, ...)
 *     sequencing/sequencing/main.py:316-322    U*psi / U*rho*U^dag between Sequence segments
 *     sequencing/sequencing/main.py:333-335    qutip.expect(op, states)
 * The reference has no FFI of its own (pure Python calling QuTiP), so each entry point cites
 * the Python call it replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All matrices are complex128 stored as
 *    interleaved (re, im) doubles, row-major.
 *  - unless SEQ_FLAG_HOST_POINTERS is set, every data pointer is a DEVICE pointer owned by the
 *    caller; the engine never frees or retains them.  Work is enqueued on the engine's stream;
 *    the caller synchronises that stream before reading outputs.
 *  - with SEQ_FLAG_HOST_POINTERS every data pointer is a HOST pointer; the engine stages the
 *    inputs to the device, solves, copies the outputs back and synchronises before returning.
 *  - every function returns 0 (SEQ_OK) on success, a negative SEQ_ERR_* code otherwise;
 *    seq_engine_last_error() gives the message.  Per-trajectory integration failures are
 *    reported in `status[b]` (the call itself still returns SEQ_OK): the Python host raises the
 *    same-meaning exception QuTiP raises ("ODE integration error ... increase nsteps").
 *  - one handle per device/stream; calls on one handle are not concurrent.
 */
#ifndef SEQ_ENGINE_H
#define SEQ_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEQ_ENGINE_ABI_VERSION 3

/* return codes */
#define SEQ_OK 0
#define SEQ_ERR_INVALID (-1)      /* malformed descriptor                        */
#define SEQ_ERR_CUDA (-2)         /* CUDA runtime error (message in last_error)  */
#define SEQ_ERR_UNSUPPORTED (-3)  /* shape / option outside what the kernels cover */
#define SEQ_ERR_NOMEM (-4)

/* per-trajectory status words */
#define SEQ_STATUS_OK 0
#define SEQ_STATUS_MAX_STEPS 1    /* more than `nsteps` internal steps in one output interval */
#define SEQ_STATUS_NONFINITE 2    /* NaN/Inf in the state or error estimate                    */
#define SEQ_STATUS_STEP_UNDERFLOW 3

/* flags */
#define SEQ_FLAG_KET 0x1u             /* state is a ket [N]; Schroedinger evolution (qutip.sesolve) */
#define SEQ_FLAG_NORMALIZE 0x2u       /* kets: renormalise at every output (Options.normalize_output) */
#define SEQ_FLAG_STORE_STATES 0x4u    /* write every output state to `states` (Options.store_states) */
#define SEQ_FLAG_EXPECT_COMPLEX 0x8u  /* `expect` is complex128 (non-Hermitian e_ops)               */
#define SEQ_FLAG_HOST_POINTERS 0x10u  /* all data pointers are host pointers (end-to-end call)      */
#define SEQ_FLAG_NO_ORDER_SWITCH 0x20u /* always step with the Dormand-Prince 5(4) pair (default: the
                                          controller may drop to a 4-stage RK4(3) pair while steps
                                          are capped by max_step and the error estimate is tiny)   */

/* kernel selection (0 = let the engine choose by N and by the sparsity of the operators) */
#define SEQ_METHOD_AUTO 0
#define SEQ_METHOD_GENERIC 1   /* FP64-FMA, one CTA per trajectory, any N                    */
#define SEQ_METHOD_DMMA 2      /* FP64 tensor-core (DMMA.8x8x4) tiles, rho resident in smem  */
#define SEQ_METHOD_SMALL 3     /* N <= 8: several trajectories per CTA, state in registers   */
#define SEQ_METHOD_DMMA_STREAM 4 /* DMMA with operands streamed through smem (N too big to stay resident) */
#define SEQ_METHOD_LARGE 5     /* one big system (N > 96): every SM works on ONE right-hand side through a
                                  stream-K DMMA ZGEMM; optionally rho row-sharded over the GPUs of the box */

#define SEQ_METHOD_SPARSE 6    /* row-sparse operators (the reference's Kronecker-embedded ladder / number operators,
                                  modes.py:125-207): ELL operator rows, FP64 FMA, one CTA per trajectory with the
                                  upper triangle of rho in registers.  AUTO picks it when every operator row holds
                                  only a few non-zeros; dense operators go to the tensor-core kernels */
#define SEQ_METHOD_DIAG 7      /* operators that are sums of a few shifted diagonals (every Kronecker-embedded ladder /
                                  number operator of the reference is): as SEQ_METHOD_SPARSE but every term is "own
                                  element + constant offset", no index tables.  AUTO prefers it over SPARSE */

typedef struct seq_engine seq_engine;

/*
 * One batched solve.  B trajectories share the operators and differ in their coefficient
 * tables and initial states: one trajectory = one (sequence x initial state x sweep value),
 * i.e. one `seq.run()` of the reference's sweep loops (calibration.py:120-125, 452-466).
 *
 *   H(t)      = H0 + sum_k coeff[b,k](t) * Hk[k]                      (basic.py:225-239)
 *   d rho/dt  = -i[H, rho] + sum_j (L_j rho L_j^dag - 1/2 {L_j^dag L_j, rho})
 *               with L_j in Lc (constant) and sqrt(rate[b,j](t)) * Ltd[j] (CTerm, basic.py:241-247)
 *   d psi/dt  = -i H(t) psi                                           (SEQ_FLAG_KET)
 *
 * coeff / rate tables are sampled on the uniform grid t0 + i*dt, i < n_t, and interpolated by
 * the natural cubic spline QuTiP 4.x applies to array coefficients; `rate` holds conj(c)*c of
 * the collapse coefficient at the knots.  Outputs are produced at the samples out_idx[0..n_out)
 * (ascending; integration starts at out_idx[0] and stops at out_idx[n_out-1]).
 */
typedef struct seq_solve_desc {
  uint32_t size;   /* = sizeof(seq_solve_desc) */
  uint32_t flags;  /* SEQ_FLAG_* */
  int32_t N;       /* Hilbert-space dimension */
  int32_t B;       /* trajectories */
  int32_t n_t;     /* samples per coefficient table */
  int32_t n_ch;    /* time-dependent Hamiltonian channels */
  int32_t n_c;     /* constant collapse operators */
  int32_t n_ctd;   /* time-dependent collapse operators */
  int32_t n_e;     /* expectation operators */
  int32_t n_out;   /* output samples */
  int32_t nsteps;  /* max internal steps per output interval (Options.nsteps, default 1000) */
  int32_t method;  /* SEQ_METHOD_* */
  double t0, dt;   /* coefficient grid */
  double atol, rtol;                  /* Options.atol / rtol (1e-8 / 1e-6) */
  double max_step, first_step, min_step; /* Options.*; 0 = unset */

  const void* H0;             /* c128 [N,N], or [B,N,N] when H0_batch_stride != 0 */
  int64_t H0_batch_stride;    /* complex elements between trajectories (0 = shared) */
  const void* Hk;             /* c128 [n_ch,N,N] */
  const double* coeff;        /* f64 [B,n_ch,n_t] */
  int64_t coeff_batch_stride; /* doubles between trajectories; 0 = one table shared by all */
  const double* coeff_scale;  /* optional f64 [B,n_ch]: per-trajectory multiplier (amplitude sweeps) */
  const void* Lc;             /* c128 [n_c,N,N] */
  const void* Ltd;            /* c128 [n_ctd,N,N] */
  const double* rate;         /* f64 [B,n_ctd,n_t] */
  int64_t rate_batch_stride;  /* doubles between trajectories; 0 = shared */
  const void* state0;         /* c128 [B,N,N] (density matrices) or [B,N] (kets) */
  const void* E;              /* c128 [n_e,N,N] */
  const int32_t* out_idx;     /* i32 [n_out] */

  void* expect;        /* f64 [B,n_e,n_out], or c128 with SEQ_FLAG_EXPECT_COMPLEX; may be NULL if n_e == 0 */
  void* final_state;   /* c128 [B,N,N] / [B,N] */
  void* states;        /* c128 [B,n_out,N,N] / [B,n_out,N]; required with SEQ_FLAG_STORE_STATES */
  int32_t* status;     /* i32 [B] */
  int32_t* stats;      /* i32 [B,4]: accepted steps, rejected steps, RHS evaluations, reserved */
} seq_solve_desc;

/* Library / ABI identification. */
int seq_engine_abi_version(void);
const char* seq_engine_build_info(void);

/* Create an engine bound to CUDA device `device` and stream `stream` (a cudaStream_t, NULL =
 * the legacy default stream).  Replaces: constructing the solver state inside qutip.mesolve. */
int seq_engine_create(int device, void* stream, seq_engine** out);
int seq_engine_destroy(seq_engine* eng);
const char* seq_engine_last_error(const seq_engine* eng);

/* Bytes of device scratch a solve of this shape needs (the engine allocates and caches it). */
int64_t seq_engine_workspace_bytes(const seq_engine* eng, const seq_solve_desc* desc);

/* The hot path.  Replaces qutip.mesolve / qutip.sesolve at basic.py:504-512. */
int seq_engine_solve(seq_engine* eng, const seq_solve_desc* desc);

/*
 * Row-sharded solve of ONE large system across `world` processes (one per GPU), config C5 of
 * BASELINE.json (transmon + two cavities, N = 1875).  Every rank passes the SAME descriptor (full
 * operators, full initial state, full-size output buffers) and gets the full result back; rank r
 * evaluates rows [r*m, (r+1)*m) of every right-hand side, m = 16*ceil(ceil(N/world)/16).  The stage
 * state is exchanged once per RHS evaluation through `all_gather`, the error norm once per step
 * through `all_reduce_sum`; both hooks are called with DEVICE pointers and must enqueue the
 * collective on the engine's stream (ncclAllGather / ncclAllReduce on that stream, or
 * torch.distributed on the stream the engine was created with).  `comm == NULL` or world == 1: one GPU.
 * Replaces: the single qutip.mesolve call of basic.py:504-512 for a system too large for one SM.
 */
typedef struct seq_comm {
  int32_t rank, world;
  void* ctx;
  /* recv[r*count .. (r+1)*count) <- rank r's send[0..count)   (doubles) */
  int (*all_gather)(void* ctx, const double* send, double* recv, int64_t count);
  /* buf[0..count) <- sum over ranks, in place (doubles) */
  int (*all_reduce_sum)(void* ctx, double* buf, int64_t count);
} seq_comm;
int seq_engine_solve_sharded(seq_engine* eng, const seq_solve_desc* desc, const seq_comm* comm);

/* Native NCCL communicator for seq_engine_solve_sharded (no hooks, no host callback per collective): the engine opens
 * libnccl.so.2 at run time (the library the host process already carries - no link-time dependency) and issues
 * ncclAllGather / ncclAllReduce on its own stream.  Rank 0 obtains an id (128 bytes), the host broadcasts it by whatever
 * means it has (torch.distributed, MPI, a file), every rank calls _init with its rank; afterwards
 * seq_engine_solve_sharded(eng, desc, NULL) shards over that communicator.  _destroy releases it. */
int seq_engine_nccl_unique_id(void* id_out_128_bytes);
int seq_engine_nccl_init(seq_engine* eng, const void* id_128_bytes, int32_t rank, int32_t world);
int seq_engine_nccl_destroy(seq_engine* eng);

/* Method the engine would pick for this descriptor (SEQ_METHOD_*), without running it. */
int seq_engine_select_method(const seq_engine* eng, const seq_solve_desc* desc);

/* Method the last seq_engine_solve on this handle actually ran (SEQ_METHOD_*): AUTO looks at the
 * operators' sparsity on the device, which seq_engine_select_method (shape only) cannot. */
int seq_engine_last_method(const seq_engine* eng);
/* One-line description of the kernel variant the last solve ran (launch shape of the diagonal kernel,
 * "large/diagonals" vs "large/zgemm", ...); empty when the method has a single variant. */
const char* seq_engine_last_variant(const seq_engine* eng);
/* FP64 pipe instructions (multiply-adds; a lone multiply counts as one) ONE right-hand-side evaluation of ONE trajectory
 * executes in the structure-exploiting kernel the last solve ran, counted by its host plan from the term lists
 * (zero terms a thread skips statically are not counted); 0 when the kernel does not report it.  bench.py turns it
 * into the executed fraction of the FP64-FMA peak. */
int64_t seq_engine_last_fma_per_rhs(const seq_engine* eng);
/* FP64 pipe instructions of one accepted 4(3) step of one trajectory OUTSIDE its right-hand sides (stage combinations, error
 * estimate incl. the IEEE square root / division sequences), counted the same way; 0 when not reported. */
int64_t seq_engine_last_step_extra_ops(const seq_engine* eng);

/* Number of kernel launches issued by the last seq_engine_solve on this handle, and the
 * duration in milliseconds of its integration kernel measured with CUDA events on the
 * engine's stream (valid after the stream is synchronised). */
int seq_engine_last_launch_count(const seq_engine* eng);
float seq_engine_last_kernel_ms(seq_engine* eng);

/* Natural-cubic-spline second derivatives M[s,:] of n_series tables y[s,0..n_t) on a uniform
 * grid (device pointers).  Replaces qutip's _prep_cubic_spline (QobjEvo array coefficients). */
int seq_engine_spline_prep(seq_engine* eng, const double* y, double* M, int64_t n_series, int32_t n_t, double dt);

/*
 * On-device waveform synthesis: the coefficient tables of a sweep are generated on
 * the GPU from per-trajectory pulse parameters instead of being built sample by sample in Python and copied over.
 * One descriptor = one chopped Gaussian(-DRAG) pulse of one trajectory:
 *     c(t_k) = amp (I_k + i Q_k) exp(-2 pi i detune tau_k) exp(i phase),          k = 0 .. n-1, n = int(chop sigma / dt)
 *     I_k = (g_k - g_0) / (1 - g_0),  g_k = exp(-t_k^2 / 2 sigma^2),  t_k = linspace(-chop sigma / 2, chop sigma / 2, n)[k]
 *     Q_k = drag t_k g_k / (4 sigma^2 (1 - g_0)),                      tau_k = linspace(0, n dt, n)[k]
 * and sign_re Re c / sign_im Im c are ADDED to coeff[traj, ch_re / ch_im, start + k] (channel -1: skipped).
 * Replaces: pulses.py:31-117 (array_pulse, gaussian_wave, gaussian_deriv_wave), pulses.py:262-293
 * (Pulse.__call__) and the quadrature split of modes.py:719-753 (Qubit.rotate: x <- Re, y <- Im) and
 * modes.py:864-897 (Cavity.displace: x <- -Im, y <- Re) for sweeps over amp / drag / detune / phase.
 * `descs` is a HOST array; `coeff` a DEVICE buffer f64 [B, n_ch, n_t] which the call zeroes first unless
 * `accumulate` is non-zero.
 */
typedef struct seq_pulse_desc {
  int32_t traj, ch_re, ch_im, start;
  double sign_re, sign_im;
  double sigma, chop, dt;
  double amp, drag, detune, phase;
} seq_pulse_desc;
int seq_engine_synth_gaussian(seq_engine* eng, const seq_pulse_desc* descs, int64_t n_desc, double* coeff, int32_t B,
                              int32_t n_ch, int32_t n_t, int accumulate);

/* states[b] <- U states[b] U^dag (density matrices) or U states[b] (kets); U is c128 [N,N]
 * shared by the batch (device pointers).  Replaces main.py:316-322. */
int seq_engine_apply_unitary(seq_engine* eng, const void* U, void* states, int32_t N, int32_t B, int is_ket);

/* out[b,e] = Tr(E[e] states[b]) or <psi_b|E[e]|psi_b>, complex128 (device pointers).
 * Replaces qutip.expect(op, states) at main.py:333-335. */
int seq_engine_expect(seq_engine* eng, const void* E, const void* states, void* out, int32_t N, int32_t B,
                      int32_t n_e, int is_ket);

#ifdef __cplusplus
}
#endif
#endif /* SEQ_ENGINE_H */
