// SEQ_METHOD_DIAG: density matrices whose operators are sums of a few (shifted) diagonals - exactly
// what the reference builds: Mode.a / a^dag / n and System.H0 are Kronecker products of identity,
// ladder and number operators (modes.py:125-207, 292-305; system.py:399-445), so in the product basis
// every operator is a handful of diagonals A[i, i+o] = v_o[i] (a (x) 1 -> one diagonal at the mode's
// stride, n -> the main diagonal, an exchange coupling a_q a_c^dag -> one diagonal at stride_q - stride_c).
//
//     (G rho)_ij       = sum_d v_d[i]        rho[i + o_d, j]
//     (rho G^dag)_ij   = sum_d conj(v_d[j])  rho[i, j + o_d]
//     (L rho L^dag)_ij = sum_{a,b} lam_a[i] conj(lam_b[j]) rho[i + o_a, j + o_b]
//
// so every term is "own element + constant offset" in the row-major stage matrix: no index tables, no
// address arithmetic per term (the host precomputes every offset in bytes).  Everything static that acts on
// the own element (G0's main diagonal: detunings, Kerr and dispersive shifts; constant-rate dephasing pairs)
// is folded into ONE complex weight per element at kernel start.  A coefficient is zero wherever its source
// would fall outside the matrix, so an out-of-range address only has to be readable and finite: Y sits
// between zero guard bands when they fit in shared memory (GUARD), otherwise the address is clamped.
//
// Work decomposition ("band" ownership).  rho is Hermitian, so of every pair (i,j), (j,i) only one element
// is integrated: row i owns the cyclic band j = i, i+1, .., i + N/2 (mod N) - every row the same length
// L = N/2 + 1, unlike the rows of a triangle.  One quarter-warp (8 lanes) owns one row, lane q the elements
// d = q, q + 8, ..: every shared-memory gather of a quarter-warp is 8 consecutive complex numbers (one
// conflict-free 128-byte wavefront) and the row coefficient v_d[i] is loaded once per thread.  y, the RK stage
// derivatives and the accumulators live in registers; the full stage matrix Y is in shared memory (owners
// write Y_ij and its mirror Y_ji), double-buffered when it fits; one barrier per right-hand side.
//   N <= 48: one CTA per trajectory.   48 < N <= 94: a cluster of two CTAs per trajectory - each owns every
//   other row, writes its elements into BOTH CTAs' shared memory (DSMEM stores) and the per-RHS barrier is a
//   cluster barrier; the error norm is exchanged the same way, the scalar control runs redundantly.
// Step control (the shared StepCtl: identical decisions to every other kernel), the spline values of all stage
// times of the next step and the output bookkeeping run on warp 0 only and reach the other warps through a
// small control block - the scalar pow()/ceil()/division chains would otherwise be re-executed by every warp.
// The stage loop is unrolled at compile time per embedded pair (static register indices, tableau entries as
// immediates).  Operators that are row-sparse but not diagonal-structured take the ELL kernel
// (SEQ_METHOD_SPARSE), dense ones the tensor-core kernels.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace seq {

#define DIAG_MAX_GD 16      // diagonals of G (union over G0 and every Gk)
#define DIAG_MAX_TAB 40     // diagonals of all collapse operators together
#define DIAG_MAX_LT 48      // (diagonal a, diagonal b) pairs of all collapse operators together
#define DIAG_E_SMEM 512     // expectation-operator non-zeros kept in shared memory

struct DiagOps {
  // --- G: main diagonal (g0_tab < 0: absent) and off-diagonal terms; all offsets in BYTES
  int nD, g0_tab, nGo;
  int go_tab[DIAG_MAX_GD];    // table of diagonal d inside a Gt buffer
  int go_offL[DIAG_MAX_GD];   // o_d * ld * 16: source of the left term  rho[i + o_d, j]
  int go_offR[DIAG_MAX_GD];   // o_d * 16:      source of the right term rho[i, j + o_d]
  // --- collapse pair terms on the own element (constant rate: folded into a static weight at kernel start;
  // time-dependent rate: evaluated per RHS) / on a gathered element
  int nLs, nLo, nLg;
  int ls_a[DIAG_MAX_LT], ls_b[DIAG_MAX_LT];
  int lo_a[DIAG_MAX_LT], lo_b[DIAG_MAX_LT], lo_r[DIAG_MAX_LT];   // table a, table b (bytes inside lam); rate index
  int lg_a[DIAG_MAX_LT], lg_b[DIAG_MAX_LT], lg_r[DIAG_MAX_LT], lg_off[DIAG_MAX_LT];
  int nT;                     // collapse diagonals (tables)
  int n_el, ld, ybufs, guard; // owned elements, leading dimension, Y buffers, guard band (elements)
  int stage_g, stage_e, stage_tab, nnzE;   // what is staged in shared memory
  int lreal;                  // every collapse diagonal is real
  int nct;                    // compute threads (controller-warp variant: blockDim has the controller's warp group on top)
  int wsm;                    // bytes of k4 .. k6 kept in shared memory (registers + controller warp)
  const cplx* g0;             // [nD][N]
  const cplx* gk;             // [n_g][nD][N]
  const cplx* lam;            // [nT][N]
  const cplx* gdiag;          // [N]: static main diagonal of G0 (folded into the own-element weight)
  const int* e_off;           // COO lists of the expectation operators (layout shared with the ELL kernel)
  const unsigned* e_ij;
  const cplx* e_val;
};

// ---------------------------------------------------------------------------------------------------
// preparation
// ---------------------------------------------------------------------------------------------------
// flags[op][o + N - 1] = 1 if operator `op` (0: union of G0, Gk; 1..n_l: L_l; 1+n_l: union of Gk) has a non-zero on diagonal o
__global__ void diag_flags_kernel(const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                  const cplx* __restrict__ L, int N, int* __restrict__ flags) {
  const int op = blockIdx.y;
  const long long N2 = (long long)N * N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N2; t += (long long)gridDim.x * blockDim.x) {
    bool nz;
    if (op == 0 || op == gridDim.y - 1) {      // 0: G0 and every Gk; last: the time-dependent part (Gk) only
      cplx v = (op == 0) ? G0[t] : cmake(0, 0);
      nz = (v.x != 0.0 || v.y != 0.0);
      for (int m = 0; m < n_g && !nz; m++) { cplx w = Gk[m * N2 + t]; nz = (w.x != 0.0 || w.y != 0.0); }
    } else {
      cplx v = L[(long long)(op - 1) * N2 + t];
      nz = (v.x != 0.0 || v.y != 0.0);
    }
    if (nz) {
      int i = (int)(t / N), j = (int)(t - (long long)i * N);
      flags[(long long)op * (2 * N - 1) + (j - i + N - 1)] = 1;
    }
  }
}

struct DiagFill {
  int nD, nT;
  int goff[DIAG_MAX_GD];
  short t_op[DIAG_MAX_TAB];   // collapse operator of table t
  int t_off[DIAG_MAX_TAB];    // its diagonal offset
};

// grid.x over rows, grid.y = nD + nT: extract one diagonal
__global__ void diag_fill_kernel(const __grid_constant__ DiagFill f, const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                 const cplx* __restrict__ L, int N, cplx* __restrict__ g0, cplx* __restrict__ gk,
                                 cplx* __restrict__ lam, cplx* __restrict__ gdiag, int* __restrict__ any_imag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
  if (i >= N) return;
  const long long N2 = (long long)N * N;
  if (d == 0) gdiag[i] = G0[(long long)i * N + i];
  if (d < f.nD) {
    const int j = i + f.goff[d];
    const bool in = j >= 0 && j < N && f.goff[d] != 0;     // the static main diagonal lives in gdiag
    g0[(long long)d * N + i] = in ? G0[(long long)i * N + j] : cmake(0, 0);
    const bool ink = j >= 0 && j < N;
    for (int m = 0; m < n_g; m++) gk[((long long)m * f.nD + d) * N + i] = ink ? Gk[m * N2 + (long long)i * N + j] : cmake(0, 0);
  } else {
    const int t = d - f.nD;
    const int j = i + f.t_off[t];
    cplx v = (j >= 0 && j < N) ? L[(long long)f.t_op[t] * N2 + (long long)i * N + j] : cmake(0, 0);
    lam[(long long)t * N + i] = v;
    if (v.y != 0.0) *any_imag = 1;
  }
}

// ---------------------------------------------------------------------------------------------------
// the integration kernel
// ---------------------------------------------------------------------------------------------------
// control block written by warp 0, read by everybody after the step barrier
struct DiagCtrl {
  double htry;
  int cmd;            // 0: take a step, 1: done
  int pair, st0;      // embedded pair and first stage to evaluate (1: k1 is known)
  int accepted;       // commit the step that was just evaluated
  int emit_lo, emit_hi;   // outputs [emit_lo, emit_hi) are produced from the committed state
  int cvbuf;          // which buffer of spline values belongs to this step (controller-warp variant)
  int valid;          // prediction slot: a prediction was made
};

// controller state, kept in shared memory between calls (warp 0 only)
struct DiagCtlState {
  StepCtl ctl;
  double htry, target;
  int o, nst, landing, capped, dbg;
  int nstA;           // nst right after the accept/reject decision (before the output loop resets it)
};

struct DiagLayout {
  int off_Y, ystride, ybytes, off_Gt, gtbytes, off_lam, off_g, off_e, off_tab, off_cv, cvstride, off_red, off_ctrl, off_w, total;
};
__host__ __device__ inline DiagLayout diag_layout(int N, int n_g, int n_t, const DiagOps& o) {
  DiagLayout L;
  const int c = (int)sizeof(cplx);
  const int gbytes = o.guard * c;
  L.ybytes = N * o.ld * c;
  L.ystride = L.ybytes + gbytes;
  L.off_Y = gbytes;
  L.off_Gt = gbytes + o.ybufs * L.ystride;
  L.gtbytes = o.nD * N * c;
  L.off_lam = L.off_Gt + 2 * L.gtbytes;
  L.off_g = L.off_lam + (o.nT > 0 ? o.nT : 1) * N * c;
  L.off_e = L.off_g + (o.stage_g ? (1 + n_g) * L.gtbytes : 0);
  L.off_tab = L.off_e + (o.stage_e ? (o.nnzE > 0 ? o.nnzE : 1) * 32 : 0);   // 16 B value + 4 B index, padded
  L.off_cv = L.off_tab + (o.stage_tab ? n_g * n_t * (int)sizeof(double2) : 0);
  L.cvstride = 7 * (n_g > 0 ? n_g : 1);
  L.off_red = L.off_cv + 2 * L.cvstride * (int)sizeof(double);
  L.off_ctrl = L.off_red + 128 * (int)sizeof(double);
  L.off_w = L.off_ctrl + 2 * 64 + 2 * 192;   // DiagCtrl[2] (actual, predicted) + DiagCtlState[2]
  L.total = L.off_w + o.wsm;
  return L;
}

__device__ __forceinline__ double diag_block_sum(double v, double* red, int& tog) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[tog * 32 + w] = v;
  __syncthreads();
  double s = 0.0;
  for (int q = 0; q < nw; q++) s += red[tog * 32 + q];
  tog ^= 1;
  return s;
}

// Warp 0, once per step: judge the step that was just evaluated (update != 0: `es` is the summed weighted
// squared error), then decide what happens next - which outputs the committed state produces, and either one
// more step (with the spline values of all its stage times) or the end.  Not inlined: its scalar state lives
// in shared memory and must not occupy registers of the element loops.
__device__ __noinline__ void diag_controller(const SolveParams& p, DiagCtlState* S, DiagCtrl* ctrl, double* cv,
                                             const double2* tabc, const double2* tabr, const double* cscale,
                                             int update, int nrhs, double es, int cvbuf) {
  const int lane = threadIdx.x & 31, n_g = p.n_g;
  PROF_DECL
  StepCtl ctl = S->ctl;
  double c_htry = S->htry, c_target = S->target;
  int c_o = S->o, c_nst = S->nst;
  bool c_landing = S->landing != 0, c_capped = S->capped != 0;
  __syncwarp();
  PROF(8);
  bool accepted = false;
  if (update) {
    ctl.nrhs += nrhs;
    ctl.have_k1 = true;
    const double err = sqrt(es / ((double)p.N * (double)p.N));
#ifdef SEQ_PROFILE_REGIONS
    if (lane == 0) { if (err <= 2.9e-6) S->dbg += 1; if (err <= 2.9e-7) S->dbg += 1 << 16; }   // instrumentation only
#endif
    accepted = ctl_update(ctl, p, err, c_htry, c_landing, c_capped, c_target, c_nst);
  }
  const int nstA = c_nst;
  PROF(9);
  const int emit_lo = c_o;
  int cmd = 1;
  while (ctl.status == SEQ_STATUS_OK && c_o < p.n_out) {
    c_target = p.t0 + p.dt * (double)p.out_idx[c_o];
    if (c_o > 0 && ctl_next(ctl, p, c_target, c_htry, c_landing, c_capped)) { cmd = 0; break; }
    c_o++;           // output c_o is reached by the committed state
    c_nst = 0;
  }
  PROF(10);
  const int emit_hi = (ctl.status == SEQ_STATUS_OK) ? c_o : emit_lo;
  const int pair = ctl.pair;
  if (cmd == 0) {
    for (int q = lane; q < RK_STAGES[pair] * n_g; q += 32) {
      const int st = q / n_g, m = q - st * n_g;
      const double ts = ctl.t + RK_C[pair][st] * c_htry;
      double v;
      if (m < p.n_ch) {
        v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, ts);
        if (cscale) v *= cscale[m];
      } else {
        v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, ts);
      }
      cv[q] = v;
    }
  }
  PROF(11);
  if (lane == 0) {
    ctrl->htry = c_htry;
    ctrl->cmd = cmd;
    ctrl->pair = pair;
    ctrl->st0 = ctl.have_k1 ? 1 : 0;
    ctrl->accepted = accepted ? 1 : 0;
    ctrl->emit_lo = emit_lo;
    ctrl->emit_hi = emit_hi;
    ctrl->cvbuf = cvbuf;
    ctrl->valid = 0;
    S->nstA = nstA;
    S->ctl = ctl;
    S->htry = c_htry; S->target = c_target;
    S->o = c_o; S->nst = c_nst; S->landing = c_landing ? 1 : 0; S->capped = c_capped ? 1 : 0;
  }
  __syncwarp();
  PROF(12);
}

// Controller-warp variant: was the prediction `Sp` (the controller run on a copy of `S` with the previous step's
// error estimate while the step was still being evaluated) the decision the real error estimate leads to?  The
// accept/reject logic is re-run on the real estimate; every field of the controller except the step-size
// proposal h must agree, and h must not bind in either (both proposals at or above max_step: the next trial
// step, its landing / capped flags and therefore the spline values are then identical).  On success the
// prediction becomes the state, with the real h.
__device__ __noinline__ bool diag_verify(const SolveParams& p, const DiagCtlState* S, DiagCtlState* Sp, int nrhs, double es) {
  StepCtl ct = S->ctl;
  int nst = S->nst;
  ct.nrhs += nrhs;
  ct.have_k1 = true;
  const double err = sqrt(es / ((double)p.N * (double)p.N));
  const bool acc = ctl_update(ct, p, err, S->htry, S->landing != 0, S->capped != 0, S->target, nst);
  const StepCtl cp = Sp->ctl;
  const bool same = acc && p.max_step > 0 && ct.h >= p.max_step && cp.h >= p.max_step && ct.t == cp.t && ct.pair == cp.pair &&
                    ct.status == cp.status && ct.nacc == cp.nacc && ct.nrej == cp.nrej && ct.nrhs == cp.nrhs &&
                    ct.probe_wait == cp.probe_wait && ct.probe_backoff == cp.probe_backoff && ct.have_k1 == cp.have_k1 &&
                    ct.prev_rejected == cp.prev_rejected && nst == Sp->nstA;
  __syncwarp();
  if (same && (threadIdx.x & 31) == 0) Sp->ctl.h = ct.h;
  __syncwarp();
  return same;
}

// Sufficient condition for diag_verify to confirm a prediction, cheap enough for the critical path (no square roots,
// no divisions).  The step being judged ran on the 4(3) pair at a capped size and was not preceded by a rejection;
// es <= (0.49 N)^2 means err <= 0.49 < SEQ_STAY_ERR with a margin far above the rounding of the exact expression.
// Then ctl_update accepts, keeps the pair (probe_backoff = 1), and proposes h = htry * min(10, 0.9 err^(-1/4)) >=
// 1.07 htry >= max_step (third line).  Two error sums that both pass therefore lead to the same controller state up to
// h, and h does not bind: exactly what diag_verify checks.
__device__ __forceinline__ bool diag_fast_state_ok(const SolveParams& p, const DiagCtlState* S) {
  return S->ctl.pair == 1 && S->capped != 0 && !S->ctl.prev_rejected && S->ctl.status == SEQ_STATUS_OK && p.stay_err <= 0.5 &&
         p.max_step > 0 && S->nst + 1 <= p.nsteps && S->htry * 1.05 >= p.max_step &&
         p.max_step >= 1e-10 * fmax(1.0, fabs(S->target) + fabs(S->ctl.t) + S->htry);
}
__device__ __forceinline__ bool diag_fast_es_ok(double es, double es_max) { return es >= 0.0 && es <= es_max; }

// named barriers of the controller-warp variant (0 is __syncthreads)
#define DIAG_BAR_STAGE 1   // compute warps only
#define DIAG_BAR_LAST 2    // last stage of a step: compute warps + controller (publishes the prediction)
#define DIAG_BAR_ERR 3     // compute warps arrive with their error partial sums, the controller waits
#define DIAG_BAR_VERDICT 4 // the controller arrives with the verdict / the actual control block, compute warps wait
#define DIAG_BAR_ALL 5     // compute warps + controller (set-up)
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// (producer side of the PTX ISA's bar.arrive / bar.sync example: shared-memory stores before the arrive are visible to
// the threads that bar.sync on the same barrier - no fence in between)
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// threads per CTA by (elements per thread, stage vectors in registers): 65536 registers / threads, allocated per 4 warps
// Inner loops over a thread's elements run in chunks of at most 3 with a compiler barrier in between: the loads
// of all EPT elements of one term would otherwise be hoisted together (EPT = 6: 72 registers of operands in
// flight on top of the 130 of persistent state -> spills to local memory that miss the small L1)
#define DIAG_CHUNK 3
#define DIAG_FOR_CHUNKED(e)                                                                         \
  _Pragma("unroll") for (int e##0 = 0; e##0 < EPT; e##0 += DIAG_CHUNK)                              \
    _Pragma("unroll") for (int e = (diag_chunk_fence(e##0), e##0); e < (e##0 + DIAG_CHUNK < EPT ? e##0 + DIAG_CHUNK : EPT); e++)
__device__ __forceinline__ void diag_chunk_fence(int e0) { if (e0 > 0) asm volatile("" ::: "memory"); }

#define DIAG_MAX_EPT 6
#define DIAG_MAXT 384       // 168 registers per thread

namespace cg = cooperative_groups;

#define DIAG_MAXT_STREAM 736   // stage vectors streamed through L2: 80 registers per thread

// CW: one extra warp (the last) carries the step controller and nothing else.  While the compute warps evaluate a
// step it PREDICTS the next control block (assuming the error estimate of the previous step; spline values of the next stage times
// included); once the error partial sums are in, it only re-runs the accept/reject logic (diag_verify) and
// publishes a verdict.  The compute warps do not wait for it: after handing in their partial sums they emit the
// outputs and stage the first RHS input of the predicted next step into the free Y / G(t) buffers (double-buffered
// Y only), and pick up the verdict at that stage's barrier.  A wrong prediction (rejected step, order switch, a
// binding step-size proposal) discards the staged input and takes the controller's actual block - decisions are
// bit-identical to the variant without the controller warp.  With CW the register budget is 128 (13 warps: one
// scheduler holds four), so k4 .. k6 (k5 / k6 are used by the 5(4) pair only) move to thread-private slots in shared
// memory; y, the static weights and k1 .. k3 stay in registers.
template <int EPT, int CL, bool GUARD, bool KREG, bool CW>
__global__ void __launch_bounds__((KREG ? DIAG_MAXT : DIAG_MAXT_STREAM) + (CW ? 32 : 0), 1)
solve_diag_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ DiagOps sp) {
  static_assert(!CW || CL == 1, "the controller warp variant is single-CTA");
  extern __shared__ __align__(16) unsigned char smem_diag[];
#ifdef SEQ_PROFILE_REGIONS
  const unsigned long long prof_c0 = clock64();
  unsigned long long prof_g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_g0));
#endif
  constexpr int C = (int)sizeof(cplx);
  constexpr int KR = KREG ? (CW ? 3 : 6) : 0;      // stage derivatives kept in registers
  const int N = p.N, ld = sp.ld, n_g = p.n_g;
  const int nth = CW ? sp.nct : (int)blockDim.x;                 // compute threads; CW: the controller warp comes on top
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (nth + 31) >> 5;
  const bool is_ctl = CW && warp == nw;
  const int nall = nth + (CW ? 32 : 0);                          // threads that take part from here on
  const int tid = threadIdx.x;
  const int rank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;      // CTA inside the cluster
  const int b = blockIdx.x / CL;                                // trajectory
  const DiagLayout L = diag_layout(N, n_g, p.n_t, sp);
  double* cv = (double*)(smem_diag + L.off_cv);
  double* red = (double*)(smem_diag + L.off_red);
  DiagCtrl* ctrl = (DiagCtrl*)(smem_diag + L.off_ctrl);
  unsigned char* peer = nullptr;                                // the other CTA's shared memory (same layout)
  if (CL > 1) peer = (unsigned char*)cg::this_cluster().map_shared_rank((void*)smem_diag, rank ^ 1);
  auto sync_all = [&]() { if (CL > 1) cg::this_cluster().sync(); else if (CW) bar_sync(DIAG_BAR_ALL, nall); else __syncthreads(); };

  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;

  // ---- stage shared tables: zero guard bands / Y padding, collapse diagonals, G diagonals, E, spline tables
  for (int q = tid; q < L.off_Gt / 16; q += nall) ((cplx*)smem_diag)[q] = cmake(0, 0);
  for (int q = tid; q < sp.nT * N; q += nall) ((cplx*)(smem_diag + L.off_lam))[q] = sp.lam[q];
  const int gsz = sp.nD * N;
  const cplx* g0p = sp.g0;
  const cplx* gkp = sp.gk;
  if (sp.stage_g) {
    cplx* gs = (cplx*)(smem_diag + L.off_g);
    for (int q = tid; q < gsz; q += nall) gs[q] = sp.g0[q];
    for (int q = tid; q < n_g * gsz; q += nall) gs[gsz + q] = sp.gk[q];
    g0p = gs; gkp = gs + gsz;
  }
  const unsigned* eij = sp.e_ij;
  const cplx* eval_ = sp.e_val;
  if (sp.stage_e) {
    cplx* ev_s = (cplx*)(smem_diag + L.off_e);
    unsigned* ei_s = (unsigned*)(ev_s + (sp.nnzE > 0 ? sp.nnzE : 1));
    for (int q = tid; q < sp.nnzE; q += nall) { ev_s[q] = sp.e_val[q]; ei_s[q] = sp.e_ij[q]; }
    eij = ei_s; eval_ = ev_s;
  }
  if (sp.stage_tab) {
    double2* ts = (double2*)(smem_diag + L.off_tab);
    for (int q = tid; q < p.n_ch * p.n_t; q += nall) ts[q] = tabc[q];
    for (int q = tid; q < p.n_ctd * p.n_t; q += nall) ts[p.n_ch * p.n_t + q] = tabr[q];
    tabc = ts; tabr = ts + p.n_ch * p.n_t;
  }

  // ---- owned elements: row i = (tid / 8) * CL + rank, band positions d = (tid % 8) + 8 e.  oy[e]: byte offset of
  // Y[i, j_e] inside a buffer; oi: of entry i of a diagonal table; rowb: of Y[i, 0] (so that oy[e] - rowb is the
  // offset of entry j_e of a table).  Laundered through an empty asm so that they stay in registers as computed.
  int oy[EPT], oi, rowb;
  unsigned vmask = 0, dmask = 0;      // valid elements; elements on the main diagonal
  // KREG: y, the static weights W and the stage derivatives k1..k6 live in registers (EPT <= 4); otherwise they are
  // thread-private coalesced streams in global memory (L2-resident), slots 0..5: k, 6: y, 7: W
  cplx ys[EPT], kout[EPT];
  constexpr bool WREG = KREG;            // static weights in registers (else: the shared L2 table)
  cplx y[KREG ? EPT : 1], W[WREG ? EPT : 1];
  cplx* wsm = (cplx*)(smem_diag + L.off_w);
  cplx k[KR > 0 ? KR : 1][KREG ? EPT : 1];
  // (KREG && CW: k4 .. k6 in shared memory)
  cplx* kst = KREG ? (CW ? (cplx*)(smem_diag + L.off_w) : nullptr) : p.ws + (long long)blockIdx.x * p.ws_stride;
  // the static weights are the same for every trajectory: ONE table behind the per-trajectory streams, written by
  // every CTA with identical values (keeps 1/8 of the stream footprint out of L2 / HBM)
  cplx* wtab = KREG ? nullptr : p.ws + (long long)gridDim.x * p.ws_stride;
  auto st_at = [&](int slot, int e) -> cplx& {
    return slot == 7 ? wtab[(long long)e * nth + tid] : kst[((long long)slot * EPT + e) * nth + tid];
  };
  // stage derivative j of element e: registers below KR; above, shared memory (KREG: slots j - KR) or the L2 stream (slot j)
  auto kget = [&](int j, int e) -> cplx { return (KREG && j < KR) ? k[j < KR ? j : 0][KREG ? e : 0] : st_at(KREG ? j - KR : j, e); };
  auto kset = [&](int j, int e, const cplx& v) {
    if (KREG && j < KR) k[j < KR ? j : 0][KREG ? e : 0] = v; else st_at(KREG ? j - KR : j, e) = v;
  };
  if (!is_ctl) {
    const int i_raw = (tid >> 3) * CL + rank, q8 = tid & 7;
    const bool rowok = i_raw < N;
    const int i = rowok ? i_raw : 0;
    const int Lb = N / 2 + 1;
    oi = i * C;
    rowb = i * ld * C;
    const cplx* y0 = p.state0 + (long long)b * N * N;
    const cplx gi = sp.gdiag[i];
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int d = q8 + 8 * e;
      const bool v = rowok && d < Lb && !((N & 1) == 0 && d == N / 2 && i >= N / 2);
      int j = i + (v ? d : 0);
      if (j >= N) j -= N;
      if (v) vmask |= 1u << e;
      if (v && d == 0) dmask |= 1u << e;
      oy[e] = (i * ld + j) * C;
      ys[e] = v ? y0[i * N + j] : cmake(0, 0);
      kout[e] = cmake(0, 0);
      if (KREG) {
        y[KREG ? e : 0] = ys[e];
#pragma unroll
        for (int q = 0; q < KR; q++) k[q < KR ? q : 0][KREG ? e : 0] = cmake(0, 0);
      } else {
        st_at(6, e) = ys[e];
        for (int q = 0; q < 6; q++) st_at(q, e) = cmake(0, 0);
      }
      // static own-element weight: -i (G0_ii - conj(G0_jj)) + sum over constant-rate pairs lam_a[i] conj(lam_b[j])
      const cplx gj = sp.gdiag[j];
      cplx w = cmake(gi.y + gj.y, -(gi.x - gj.x));
      for (int t = 0; t < sp.nLs; t++) {
        const cplx a = sp.lam[sp.ls_a[t] / C + i], c = sp.lam[sp.ls_b[t] / C + j];
        w.x += a.x * c.x + a.y * c.y;
        w.y += a.y * c.x - a.x * c.y;
      }
      if (WREG) W[WREG ? e : 0] = w; else if (KREG) wsm[e * nth + tid] = w; else st_at(7, e) = w;
      asm volatile("" : "+r"(oy[e]));
    }
    asm volatile("" : "+r"(oi), "+r"(rowb));
  }
  sync_all();   // guard bands zeroed (in both CTAs) before anything is written into Y

  int yb = 0, gb = 0, tog = 0, ycur = 0;
  const double* cvc = cv;      // spline values of the step being evaluated (CW: one of two buffers)
  PROF_DECL
  // owner writes Y_ij and the mirror Y_ji, into every CTA of the cluster
  auto write_Y = [&](int boff, const cplx (&v)[EPT]) {
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      if (vmask & (1u << e)) {
        const int ot = (oy[e] - rowb) * ld + oi;          // (j ld + i) 16
        *(cplx*)(smem_diag + boff + oy[e]) = v[e];
        if (CL > 1) *(cplx*)(peer + boff + oy[e]) = v[e];
        if (!(dmask & (1u << e))) {
          *(cplx*)(smem_diag + boff + ot) = cconj(v[e]);
          if (CL > 1) *(cplx*)(peer + boff + ot) = cconj(v[e]);
        }
      }
    }
  };
  auto build_Gt = [&](cplx* Gb, const double* cvs, int st) {
    for (int q = tid; q < gsz; q += nth) {
      cplx g = g0p[q];
      for (int m = 0; m < n_g; m++) {
        const cplx a = gkp[m * gsz + q];
        const double v = cvs[st * n_g + m];
        g.x = fma(v, a.x, g.x);
        g.y = fma(v, a.y, g.y);
      }
      Gb[q] = g;
    }
  };
  const int y_hi = L.ybytes - C;
  auto yaddr = [&](int a) -> int { return GUARD ? a : min(max(a, 0), y_hi); };
  const bool lreal = sp.lreal != 0;

  // kout = W ys - i (G' Y - Y G'^dag) + collapse terms, G' = G minus its static main diagonal
  auto compute_K = [&](const unsigned char* __restrict__ Yb, const unsigned char* __restrict__ Gb, int st) {
#pragma unroll
    for (int e = 0; e < EPT; e++) kout[e] = cmul(WREG ? W[WREG ? e : 0] : (KREG ? wsm[e * nth + tid] : st_at(7, e)), ys[e]);
    if (sp.g0_tab >= 0) {      // time-dependent main diagonal
      const unsigned char* gv = Gb + sp.g0_tab;
      const cplx vi = *(const cplx*)(gv + oi);
      DIAG_FOR_CHUNKED(e) {
        const cplx vj = *(const cplx*)(gv - rowb + oy[e]);
        const cplx dd = cmake(vi.y + vj.y, -(vi.x - vj.x));   // -i (v_0[i] - conj(v_0[j]))
        cfma(kout[e], dd, ys[e]);
      }
    }
#pragma unroll 1
    for (int t = 0; t < sp.nGo; t++) {
      const unsigned char* gv = Gb + sp.go_tab[t];
      const int offL = sp.go_offL[t], offR = sp.go_offR[t];
      const cplx vi = *(const cplx*)(gv + oi);
      const cplx mvi = cmake(vi.y, -vi.x);            // -i v_d[i]
      const unsigned char* gvj = gv - rowb;
      DIAG_FOR_CHUNKED(e) {
        const cplx ya = *(const cplx*)(Yb + yaddr(oy[e] + offL));
        cfma(kout[e], mvi, ya);
        const cplx vj = *(const cplx*)(gvj + oy[e]);
        const cplx yv = *(const cplx*)(Yb + yaddr(oy[e] + offR));
        // + i conj(vj) yv = (vj.y + i vj.x) (yv.x + i yv.y)
        kout[e].x = fma(vj.y, yv.x, kout[e].x); kout[e].x = fma(-vj.x, yv.y, kout[e].x);
        kout[e].y = fma(vj.y, yv.y, kout[e].y); kout[e].y = fma(vj.x, yv.x, kout[e].y);
      }
    }
    const unsigned char* lam = smem_diag + L.off_lam;
    // collapse terms on the own element with a time-dependent rate
#pragma unroll 1
    for (int t = 0; t < sp.nLo; t++) {
      const unsigned char* la = lam + sp.lo_a[t];
      const unsigned char* lbj = lam + sp.lo_b[t] - rowb;
      const double r = cvc[st * n_g + sp.lo_r[t]];
      const cplx a = seq::cscale(r, *(const cplx*)(la + oi));
      DIAG_FOR_CHUNKED(e) {
        const cplx c = *(const cplx*)(lbj + oy[e]);
        const double wx = a.x * c.x + a.y * c.y, wy = a.y * c.x - a.x * c.y;
        kout[e].x = fma(wx, ys[e].x, kout[e].x); kout[e].x = fma(-wy, ys[e].y, kout[e].x);
        kout[e].y = fma(wx, ys[e].y, kout[e].y); kout[e].y = fma(wy, ys[e].x, kout[e].y);
      }
    }
    // collapse terms that gather rho[i + o_a, j + o_b]
#pragma unroll 1
    for (int t = 0; t < sp.nLg; t++) {
      const unsigned char* la = lam + sp.lg_a[t];
      const unsigned char* lbj = lam + sp.lg_b[t] - rowb;
      const int off = sp.lg_off[t];
      const int ri = sp.lg_r[t];
      const double r = ri < 0 ? 1.0 : cvc[st * n_g + ri];
      if (lreal) {
        const double ra = r * (*(const double*)(la + oi));
        DIAG_FOR_CHUNKED(e) {
          const cplx yv = *(const cplx*)(Yb + yaddr(oy[e] + off));
          const double w = ra * (*(const double*)(lbj + oy[e]));
          kout[e].x = fma(w, yv.x, kout[e].x);
          kout[e].y = fma(w, yv.y, kout[e].y);
        }
      } else {
        const cplx a = seq::cscale(r, *(const cplx*)(la + oi));
        DIAG_FOR_CHUNKED(e) {
          const cplx yv = *(const cplx*)(Yb + yaddr(oy[e] + off));
          const cplx c = *(const cplx*)(lbj + oy[e]);
          const double wx = a.x * c.x + a.y * c.y, wy = a.y * c.x - a.x * c.y;
          kout[e].x = fma(wx, yv.x, kout[e].x); kout[e].x = fma(-wy, yv.y, kout[e].x);
          kout[e].y = fma(wx, yv.y, kout[e].y); kout[e].y = fma(wy, yv.x, kout[e].y);
        }
      }
    }
  };
  auto emit = [&](const unsigned char* __restrict__ Yb, int o) {
    // one warp per operator, starting from the LAST warp (warp 0 carries the step controller); cluster: CTA 0
    if (rank == 0) {
      for (int e = nw - 1 - warp; e < p.n_e; e += nw) {
        cplx acc = cmake(0, 0);
        const int q0 = sp.e_off[e], q1 = sp.e_off[e + 1];
        for (int q = q0 + lane; q < q1; q += 32) {
          const unsigned code = eij[q];
          cfma(acc, eval_[q], ((const cplx*)Yb)[(code >> 16) * ld + (code & 0xffffu)]);   // E_ij rho_ji
        }
        if (q1 - q0 > 1) {          // projectors (the sweeps' e_ops) have a single entry: nothing to reduce
          acc.x = warp_sum(acc.x);
          acc.y = warp_sum(acc.y);
        }
        if (lane == 0) {
          const long long off = ((long long)b * p.n_e + e) * p.n_out + o;
          if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = acc;
          else ((double*)p.expect)[off] = acc.x;
        }
      }
    }
    if (p.flags & SEQ_FLAG_STORE_STATES) {
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N * N;
      for (int q = tid + rank * nth; q < N * N; q += nth * CL) { const int i = q / N, j = q - i * N; dst[q] = ((const cplx*)Yb)[i * ld + j]; }
    }
  };

  static_assert(sizeof(DiagCtrl) <= 56 && sizeof(DiagCtlState) <= 192, "control blocks outgrew their shared-memory slots");
  DiagCtrl* ctrl_pred = (DiagCtrl*)(smem_diag + L.off_ctrl + 64);
  volatile int* verdict = (volatile int*)(smem_diag + L.off_ctrl + 56);      // in the padding behind the actual control block
  DiagCtlState* cst = (DiagCtlState*)(smem_diag + L.off_ctrl + 128);
  if (tid == (CW ? nth : 0)) {      // (CW: first lane of the controller warp)
    StepCtl c0;
    ctl_init(c0, p);
    cst->ctl = c0;
    cst->htry = 0.0; cst->target = 0.0; cst->o = 0; cst->nst = 0; cst->landing = 0; cst->capped = 0; cst->dbg = 0; cst->nstA = 0;
  }
  __syncwarp();

  if (!is_ctl) write_Y(L.off_Y, ys);
  if (CW ? is_ctl : warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 0, 0, 0.0, 0);
  sync_all();

  if (CW && is_ctl) {
    // ---------------- the controller warp ----------------
    DiagCtlState* S = cst;
    DiagCtlState* Sp = (DiagCtlState*)(smem_diag + L.off_ctrl + 128 + 192);
    int cur = 0, ctog = 0;       // spline-value buffer of the step being evaluated; toggle of the partial-sum buffer
    double es_prev = 0.0;        // the prediction assumes the error estimate of the previous step
    bool inconsistent = false;
#ifdef SEQ_PROFILE_REGIONS
    unsigned long long _pc = clock64();     // controller-warp regions 8..11 (CTA 0): predict, wait, verify, publish
#define PROFC(r) do { if (lane == 0 && blockIdx.x == 0) { unsigned long long _n = clock64(); seq_prof[r] += _n - _pc; _pc = _n; } } while (0)
#else
#define PROFC(r)
#endif
    int cmd = ctrl->cmd, nrhs = (ctrl->pair ? 5 : 7) - ctrl->st0;      // of the step being evaluated
    __syncwarp();
    while (!cmd) {
      // prediction: the controller on a copy of the state, with the error estimate of the previous step
      const bool predict = p.max_step > 0 && S->ctl.status == SEQ_STATUS_OK;
      if (predict) {
        if (lane == 0) *Sp = *S;
        __syncwarp();
        diag_controller(p, Sp, ctrl_pred, cv + (cur ^ 1) * L.cvstride, tabc, tabr, cscale, 1, nrhs, es_prev, cur ^ 1);
      }
      if (lane == 0) ctrl_pred->valid = (predict && ctrl_pred->accepted && ctrl_pred->st0 == 1) ? 1 : 0;
      __syncwarp();
      const bool valid = ctrl_pred->valid != 0;
      const int p_cmd = ctrl_pred->cmd, p_nrhs = (ctrl_pred->pair ? 5 : 7) - ctrl_pred->st0;
      // Steady state of a capped sweep: the step runs on the 4(3) pair at the capped size and both the previous and the
      // new error estimate are below the stay threshold with a margin.  Then the accept/reject logic provably ends in
      // the predicted state (accepted, pair kept, proposal above the cap - see diag_fast_state_ok), so the verdict goes out
      // right after the block sum and the exact bookkeeping (the proposal h) follows off the critical path.
      const double es_max = 0.9604 * p.stay_err * p.stay_err * (double)p.N * (double)p.N;      // err <= 0.98 stay_err
      const bool fast_prev = valid && diag_fast_state_ok(p, S) && diag_fast_es_ok(es_prev, es_max);
      __threadfence_block();
      PROFC(8);
      bar_sync(DIAG_BAR_LAST, nall);
      bar_sync(DIAG_BAR_ERR, nall);
      PROFC(9);
      double es = (lane < nw) ? red[ctog * 32 + lane] : 0.0;
      es = warp_sum(es);
      ctog ^= 1;
      const bool fast = fast_prev && diag_fast_es_ok(es, es_max);
      if (fast) {
        if (lane == 0) *verdict = 1;
        bar_arrive(DIAG_BAR_VERDICT, nall);
      }
      PROFC(10);
      es_prev = es;
      bool ok = valid && diag_verify(p, S, Sp, nrhs, es);
      if (fast && !ok) { inconsistent = true; ok = true; }     // cannot happen (the fast conditions imply diag_verify): reported below
      if (ok) {
        DiagCtlState* t = S; S = Sp; Sp = t;
        cur ^= 1;
        cmd = p_cmd; nrhs = p_nrhs;
#ifndef SEQ_PROFILE_REGIONS
        if (lane == 0) S->dbg += 1;      // stats[3]: confirmed predictions
#endif
        if (!fast) {
          if (lane == 0) *verdict = 1;
          bar_arrive(DIAG_BAR_VERDICT, nall);
        }
      } else {
        diag_controller(p, S, ctrl, cv + cur * L.cvstride, tabc, tabr, cscale, 1, nrhs, es, cur);
        __syncwarp();
        cmd = ctrl->cmd; nrhs = (ctrl->pair ? 5 : 7) - ctrl->st0;
        if (lane == 0) *verdict = 0;
        bar_arrive(DIAG_BAR_VERDICT, nall);
      }
      __syncwarp();
      PROFC(11);
    }
    if (lane == 0) {
      p.status[b] = inconsistent ? SEQ_STATUS_NONFINITE : S->ctl.status;     // a fast verdict the exact check rejected: fail loudly
      p.stats[b * 4 + 0] = S->ctl.nacc;
      p.stats[b * 4 + 1] = S->ctl.nrej;
      p.stats[b * 4 + 2] = S->ctl.nrhs;
      p.stats[b * 4 + 3] = S->dbg;
    }
    return;
  }

  auto sync_stage = [&]() { if (CW) bar_sync(DIAG_BAR_STAGE, nth); else sync_all(); };

  // One embedded-RK step with the stage loop unrolled at compile time: static register indices for the stage
  // derivatives, tableau coefficients as immediates (zero entries vanish).  Returns this thread's share of
  // the weighted squared error.  `prepped` (CW): the input of the first evaluated stage (st0 == 1) is already in
  // Y / G(t) and the barrier behind it has been passed.
  auto run_step = [&](auto tag, const int st0, const double htry, const bool prepped) -> double {
    constexpr int PAIR = decltype(tag)::value;
    constexpr int S = PAIR ? 5 : 7;
    double isc[EPT];      // error weights (1 or 2 for the mirrored element) / scale^2
#pragma unroll
    for (int st = 0; st < S; st++) {
      if (st == 0 && st0 != 0) continue;        // FSAL: k1 is known
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        double ax = 0.0, ay = 0.0;
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const double aj = rk_a_c(PAIR, st, j);
          if (j < st && aj != 0.0) {
            const cplx kv = kget(j, e);
            ax = fma(aj, kv.x, ax);
            ay = fma(aj, kv.y, ay);
          }
        }
        const cplx yv = KREG ? y[KREG ? e : 0] : st_at(6, e);
        ys[e] = (st == 0) ? yv : cmake(fma(htry, ax, yv.x), fma(htry, ay, yv.y));
      }
      if (!(CW && st == 1 && prepped)) {
        if (sp.ybufs == 2) yb ^= 1;
        gb ^= 1;
        if (CL > 1 && sp.ybufs == 1) cg::this_cluster().barrier_wait();   // every gather of the previous stage is done
        write_Y(L.off_Y + yb * L.ystride, ys);
        build_Gt((cplx*)(smem_diag + L.off_Gt + gb * L.gtbytes), cvc, st);
        PROF(1);
        if (CW && st == S - 1) bar_sync(DIAG_BAR_LAST, nall);      // the controller joins: its prediction becomes visible
        else sync_stage();
      }
      const int boff = L.off_Y + yb * L.ystride;
      unsigned char* Gb = smem_diag + L.off_Gt + gb * L.gtbytes;
      PROF(2);
      if (st == S - 1) {
        // the last stage input is the new state: the error scale (square root + division chains) is known here and
        // overlaps the gathers of the last right-hand side instead of sitting in the serial tail of the step
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          const cplx yo = KREG ? y[KREG ? e : 0] : st_at(6, e);
          const double sc = p.atol + p.rtol * sqrt(fmax(cabs2(yo), cabs2(ys[e])));
          const double wgt = (vmask & (1u << e)) ? ((dmask & (1u << e)) ? 1.0 : 2.0) : 0.0;
          isc[e] = wgt / (sc * sc);
        }
      }
      compute_K(smem_diag + boff, Gb, st);
      if (sp.ybufs == 1) {                      // all gathers done before the next stage overwrites Y
        if (CL > 1) cg::this_cluster().barrier_arrive(); else sync_stage();
      }
      if (st < S - 1) {
#pragma unroll
        for (int e = 0; e < EPT; e++) kset(st, e, kout[e]);
      }
      PROF(3);
    }
    // error estimate: h (sum_{j<S-1} e_j k_j + e_{S-1} k_S), weighted RMS over all N^2 elements
    double es = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      double ex = rk_e_c(PAIR, S - 1) * kout[e].x, ey = rk_e_c(PAIR, S - 1) * kout[e].y;
#pragma unroll
      for (int j = 0; j < S - 1; j++) {
        const double cj = rk_e_c(PAIR, j);
        if (cj != 0.0) {
          const cplx kv = kget(j, e);
          ex = fma(cj, kv.x, ex);
          ey = fma(cj, kv.y, ey);
        }
      }
      ex *= htry; ey *= htry;
      es = fma(ex * ex + ey * ey, isc[e], es);
    }
    return es;
  };
  auto commit = [&]() {
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      if (KREG) y[KREG ? e : 0] = ys[e]; else st_at(6, e) = ys[e];
      kset(0, e, kout[e]);
    }
    ycur = yb;
  };

  if (CW) {
    // ---------------- compute warps next to a controller warp ----------------
    DiagCtrl c = *ctrl;
    bool prepped = false;        // commit + outputs done, first stage input of the step staged
    while (true) {
      if (!prepped) {
        if (c.accepted) commit();
        for (int o = c.emit_lo; o < c.emit_hi; o++) emit(smem_diag + L.off_Y + ycur * L.ystride, o);
      }
      PROF(0);
      if (c.cmd) break;
      cvc = cv + c.cvbuf * L.cvstride;
      if (sp.ybufs == 1 && !prepped && c.emit_hi > c.emit_lo) sync_stage();   // the next stage overwrites the buffer emit() read
      double es;
      if (c.pair) es = run_step(IntTag<1>(), c.st0, c.htry, prepped);
      else es = run_step(IntTag<0>(), c.st0, c.htry, prepped);
      PROF(4);
      const DiagCtrl pc = *ctrl_pred;         // published before the barrier of the last stage
      es = warp_sum(es);
      if (lane == 0) red[tog * 32 + warp] = es;
      tog ^= 1;
      bar_arrive(DIAG_BAR_ERR, nall);
      PROF(5);
      prepped = false;
      const bool spec = pc.valid && sp.ybufs == 2;
      if (spec) {
        // speculation on "accepted": the outputs of the new state (rewritten later if the step turns out rejected),
        // then the first stage input of the predicted next step into the free Y / G(t) buffers
        for (int o = pc.emit_lo; o < pc.emit_hi; o++) emit(smem_diag + L.off_Y + yb * L.ystride, o);
        if (!pc.cmd) {
          const double a21 = pc.pair ? rk_a_c(1, 1, 0) : rk_a_c(0, 1, 0);
          cplx tmp[EPT];
#pragma unroll
          for (int e = 0; e < EPT; e++) {
            const double ax = fma(a21, kout[e].x, 0.0), ay = fma(a21, kout[e].y, 0.0);
            tmp[e] = cmake(fma(pc.htry, ax, ys[e].x), fma(pc.htry, ay, ys[e].y));
          }
          write_Y(L.off_Y + (yb ^ 1) * L.ystride, tmp);
          build_Gt((cplx*)(smem_diag + L.off_Gt + (gb ^ 1) * L.gtbytes), cv + pc.cvbuf * L.cvstride, 1);
        }
      }
      PROF(6);
      bar_sync(DIAG_BAR_VERDICT, nall);
      PROF(7);
      if (pc.valid && *verdict) {
        c = pc;
        if (spec) {
          commit();
          if (!c.cmd) { yb ^= 1; gb ^= 1; }
          prepped = true;
        }
      } else {
        c = *ctrl;
      }
    }
  } else {
  while (true) {
    const DiagCtrl c = *ctrl;
    if (c.accepted) commit();
    for (int o = c.emit_lo; o < c.emit_hi; o++) emit(smem_diag + L.off_Y + ycur * L.ystride, o);
    PROF(0);
    if (c.cmd) break;
    if (sp.ybufs == 1) {
      // the next stage overwrites the buffer emit() read.  Cluster: every stage starts by closing a split barrier
      // whose arrive marks "my gathers from Y are done" - open the one for the first stage of this step here
      if (CL > 1) cg::this_cluster().barrier_arrive();
      else if (c.emit_hi > c.emit_lo) __syncthreads();
    }
    double es;
    if (c.pair) es = run_step(IntTag<1>(), c.st0, c.htry, false);
    else es = run_step(IntTag<0>(), c.st0, c.htry, false);
    PROF(4);
    // block (cluster) sum of the error: every warp leaves its partial sum in every CTA, fixed summation order
    es = warp_sum(es);
    if (lane == 0) {
      red[tog * 64 + rank * 32 + warp] = es;
      if (CL > 1) ((double*)(peer + L.off_red))[tog * 64 + rank * 32 + warp] = es;
    }
    if (CL > 1 && sp.ybufs == 1) cg::this_cluster().barrier_wait();      // close the split barrier of the last stage
    sync_all();
    es = 0.0;
    for (int r = 0; r < CL; r++)
      for (int q = 0; q < nw; q++) es += red[tog * 64 + r * 32 + q];
    tog ^= 1;
    PROF(5);
    if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 1, (c.pair ? 5 : 7) - c.st0, es, 0);
    PROF(6);
    __syncthreads();
    PROF(7);
  }
  }

  cplx* fin = p.final_state + (long long)b * N * N;
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    if (vmask & (1u << e)) {
      const int i = oi / C, j = (oy[e] - rowb) / C;
      const cplx yf = KREG ? y[KREG ? e : 0] : st_at(6, e);
      fin[i * N + j] = yf;
      if (i != j) fin[j * N + i] = cconj(yf);
    }
  }
  if (!CW && tid == 0 && rank == 0) {
    p.status[b] = cst->ctl.status;
    p.stats[b * 4 + 0] = cst->ctl.nacc;
    p.stats[b * 4 + 1] = cst->ctl.nrej;
    p.stats[b * 4 + 2] = cst->ctl.nrhs;
    p.stats[b * 4 + 3] = cst->dbg;
  }
#ifdef SEQ_PROFILE_REGIONS
  if (tid == 0) {      // 13: cycles of CTA 0, 14: longest CTA, 15: nanoseconds of CTA 0
    const unsigned long long dc = clock64() - prof_c0;
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0) { seq_prof[13] += dc; seq_prof[15] += g1 - prof_g0; }
    atomicMax(&seq_prof[14], dc);
  }
#endif
  if (CL > 1) cg::this_cluster().sync();     // no CTA exits while its shared memory may still be written remotely
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct DiagPlan {
  DiagOps ops;          // offsets / term lists filled in; pointers set at launch
  DiagFill fill;
  int maxoff;
  int ept, cl, lpr, threads;    // threads: compute threads (the controller warp comes on top)
  bool kreg, guard, cw;
  size_t smem;
  long long macs;       // complex multiply-adds per element (cost model)
};

// flags: host copy of diag_flags_kernel's output, [1 + n_l][2N - 1].  Returns false if the operators are not
// diagonal-structured enough (too many diagonals) or no launch shape fits.
static inline bool diag_plan(int N, int n_g, int n_c, int n_ctd, int n_ch, int n_t, int nnzE, const int* flags, int max_smem,
                             int cl_override, bool allow_cw, DiagPlan* pl) {
  const int n_l = n_c + n_ctd, W = 2 * N - 1, c = (int)sizeof(cplx);
  if (N < 2 || N > 0x7fff) return false;
  DiagOps& o = pl->ops;
  DiagFill& f = pl->fill;
  memset(&o, 0, sizeof(o));
  memset(&f, 0, sizeof(f));
  o.ld = N | 1;
  o.nD = 0; o.g0_tab = -1; o.nGo = 0;
  int maxoff = 0;
  for (int q = 0; q < W; q++) {
    if (!flags[q]) continue;
    if (o.nD >= DIAG_MAX_GD) return false;
    const int off = q - (N - 1);
    f.goff[o.nD] = off;
    if (off == 0) {
      if (!flags[(long long)(1 + n_l) * W + q]) continue;     // static main diagonal only: folded into the own weight
      o.g0_tab = o.nD * N * c;
    } else {
      o.go_tab[o.nGo] = o.nD * N * c;
      o.go_offL[o.nGo] = off * o.ld * c;
      o.go_offR[o.nGo] = off * c;
      o.nGo++;
    }
    o.nD++;
    maxoff = abs(off) > maxoff ? abs(off) : maxoff;
  }
  if (o.nD == 0) { f.goff[0] = 0; o.nD = 1; }   // nothing but a static main diagonal: keep one (unused) table
  f.nD = o.nD;
  int nT = 0;
  o.nLs = o.nLo = o.nLg = 0;
  for (int l = 0; l < n_l; l++) {
    const int first = nT;
    for (int q = 0; q < W; q++) {
      if (!flags[(long long)(1 + l) * W + q]) continue;
      if (nT >= DIAG_MAX_TAB) return false;
      f.t_op[nT] = (short)l; f.t_off[nT] = q - (N - 1);
      maxoff = abs(f.t_off[nT]) > maxoff ? abs(f.t_off[nT]) : maxoff;
      nT++;
    }
    if (nT - first > 4) return false;
    const int rate = (l < n_c) ? -1 : (n_ch + (l - n_c));
    for (int a = first; a < nT; a++)
      for (int b2 = first; b2 < nT; b2++) {
        if (f.t_off[a] == 0 && f.t_off[b2] == 0) {
          if (o.nLo >= DIAG_MAX_LT || o.nLs >= DIAG_MAX_LT) return false;
          if (rate < 0) { o.ls_a[o.nLs] = a * N * c; o.ls_b[o.nLs] = b2 * N * c; o.nLs++; }
          else { o.lo_a[o.nLo] = a * N * c; o.lo_b[o.nLo] = b2 * N * c; o.lo_r[o.nLo] = rate; o.nLo++; }
        } else {
          if (o.nLg >= DIAG_MAX_LT) return false;
          o.lg_a[o.nLg] = a * N * c; o.lg_b[o.nLg] = b2 * N * c; o.lg_r[o.nLg] = rate;
          o.lg_off[o.nLg] = (f.t_off[a] * o.ld + f.t_off[b2]) * c;
          o.nLg++;
        }
      }
  }
  o.nT = nT; f.nT = nT;
  o.nnzE = nnzE;
  pl->maxoff = maxoff;
  pl->macs = 2LL * o.nD + o.nLs + o.nLo + o.nLg;
  o.n_el = N * (N + 1) / 2;
  // launch shape: one quarter-warp per row of the band; N <= 48 one CTA, N <= 94 a cluster of two CTAs
  const int Lb = N / 2 + 1;
  pl->lpr = 8;
  pl->ept = (Lb + 7) / 8;
  if (pl->ept > DIAG_MAX_EPT) return false;
  if (8 * N <= DIAG_MAXT) {                // N <= 48: one CTA, everything in registers
    pl->cl = 1; pl->kreg = true;
    pl->threads = (8 * N + 31) / 32 * 32;
    if (cl_override == 2 && N >= 16) { pl->cl = 2; pl->threads = (8 * ((N + 1) / 2) + 31) / 32 * 32; }
  } else if (8 * N <= DIAG_MAXT_STREAM && cl_override != 2) {   // N <= 92: one CTA, stage vectors streamed through L2
    pl->cl = 1; pl->kreg = false;
    pl->threads = (8 * N + 31) / 32 * 32;
  } else {                                 // (override) a cluster of two CTAs with register-resident stage vectors
    pl->cl = 2; pl->kreg = true;
    pl->threads = (8 * ((N + 1) / 2) + 31) / 32 * 32;
    if (pl->threads > DIAG_MAXT) return false;
  }
  // (16 lanes per row - twice the warps at 80 registers - was measured slower on C2: 9.7k vs 13.0k trajectories/s;
  // the kernel is bound by shared-memory bandwidth, not by latency)
  // shared memory, in order of value: Y (guard bands + double buffering, then clamped addresses, then a single
  // buffer), then the G diagonals, the spline tables and the expectation operators as far as they fit
  const int guard_el = maxoff * o.ld + maxoff + 1;
  const int tries[3][2] = {{2, 1}, {2, 0}, {1, 0}};
  bool ok = false;
  for (int q = 0; q < 3 && !ok; q++) {
    if (tries[q][1] && (pl->cl > 1 || !pl->kreg)) continue;   // clusters and streamed variants always clamp
    o.ybufs = tries[q][0];
    pl->guard = tries[q][1] != 0;
    o.guard = pl->guard ? guard_el : 0;
    o.stage_g = o.stage_e = o.stage_tab = 0;
    ok = diag_layout(N, n_g, n_t, o).total <= max_smem;
  }
  if (!ok) return false;
  o.stage_g = 1;
  if (diag_layout(N, n_g, n_t, o).total > max_smem) o.stage_g = 0;
  o.stage_tab = (long long)n_g * n_t * 16 < (1 << 20) ? 1 : 0;
  if (o.stage_tab && diag_layout(N, n_g, n_t, o).total > max_smem) o.stage_tab = 0;
  o.stage_e = (nnzE > 0 && nnzE <= DIAG_E_SMEM) ? 1 : 0;
  if (o.stage_e && diag_layout(N, n_g, n_t, o).total > max_smem) o.stage_e = 0;
  // a controller warp next to the compute warps, where a step is short enough for the controller to matter (N <= 48)
  pl->cw = allow_cw && pl->cl == 1 && pl->kreg;
  o.wsm = 0;
  o.nct = pl->threads;
  if (pl->cw && pl->kreg) {             // 13 warps: 128 registers per thread - k4 .. k6 move to shared memory
    o.wsm = 3 * pl->ept * pl->threads * c;      // k4, k5, k6 (k3 as well would be 1 % faster on C2 but no longer fits next to larger operator tables)
    if (diag_layout(N, n_g, n_t, o).total > max_smem) { o.wsm = 0; pl->cw = false; }
  }
  pl->smem = (size_t)diag_layout(N, n_g, n_t, o).total;
  return true;
}

static inline size_t diag_ws_elems(const DiagPlan& pl) { return pl.kreg ? 0 : (size_t)7 * pl.ept * pl.threads; }   // k1..k6, y
static inline size_t diag_ws_shared_elems(const DiagPlan& pl) { return pl.kreg ? 0 : (size_t)pl.ept * pl.threads; }  // W

struct DiagPack { size_t o_g0, o_gk, o_lam, o_gdiag, o_flag, bytes; };
static inline DiagPack diag_pack_layout(int N, int n_g, const DiagPlan& pl) {
  auto up = [](size_t x) { return (x + 15) / 16 * 16; };
  DiagPack k;
  size_t o = 0;
  k.o_g0 = o; o += (size_t)pl.ops.nD * N * sizeof(cplx);
  k.o_gk = o; o += up((size_t)(n_g > 0 ? n_g : 1) * pl.ops.nD * N * sizeof(cplx));
  k.o_lam = o; o += up((size_t)(pl.ops.nT > 0 ? pl.ops.nT : 1) * N * sizeof(cplx));
  k.o_gdiag = o; o += up((size_t)N * sizeof(cplx));
  k.o_flag = o; o += 16;
  k.bytes = o + 64;
  return k;
}

#ifndef SEQ_DIAG_NO_LAUNCH      // (single-instance compile checks define it)
template <int EPT, int CL, bool GUARD, bool KREG, bool CW>
static int launch_diag_t(const SolveParams& p, const DiagOps& sp, const DiagPlan& pl, cudaStream_t st) {
  auto kern = solve_diag_kernel<EPT, CL, GUARD, KREG, CW>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return SEQ_ERR_CUDA;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)p.B * CL, 1, 1);
  cfg.blockDim = dim3((unsigned)pl.threads + (CW ? 32 : 0), 1, 1);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p, sp) == cudaSuccess ? SEQ_OK : SEQ_ERR_CUDA;
}

template <bool CW>
static inline int launch_diag_cl1(const SolveParams& p, const DiagOps& sp, const DiagPlan& pl, cudaStream_t st) {
  if (!pl.kreg) {
    if (CW) return SEQ_ERR_UNSUPPORTED;
    switch (pl.ept) {
      case 4: return launch_diag_t<4, 1, false, false, false>(p, sp, pl, st);
      case 5: return launch_diag_t<5, 1, false, false, false>(p, sp, pl, st);
      case 6: return launch_diag_t<6, 1, false, false, false>(p, sp, pl, st);
    }
  } else if (pl.guard) {
    switch (pl.ept) {
      case 1: return launch_diag_t<1, 1, true, true, CW>(p, sp, pl, st);
      case 2: return launch_diag_t<2, 1, true, true, CW>(p, sp, pl, st);
      case 3: return launch_diag_t<3, 1, true, true, CW>(p, sp, pl, st);
      case 4: return launch_diag_t<4, 1, true, true, CW>(p, sp, pl, st);
    }
  } else {
    switch (pl.ept) {
      case 1: return launch_diag_t<1, 1, false, true, CW>(p, sp, pl, st);
      case 2: return launch_diag_t<2, 1, false, true, CW>(p, sp, pl, st);
      case 3: return launch_diag_t<3, 1, false, true, CW>(p, sp, pl, st);
      case 4: return launch_diag_t<4, 1, false, true, CW>(p, sp, pl, st);
    }
  }
  return SEQ_ERR_UNSUPPORTED;
}

static inline int launch_diag_kernel(const SolveParams& p, const DiagOps& sp, const DiagPlan& pl, cudaStream_t st) {
  if (pl.cl == 1) return pl.cw ? launch_diag_cl1<true>(p, sp, pl, st) : launch_diag_cl1<false>(p, sp, pl, st);
  switch (pl.ept) {
    case 1: return launch_diag_t<1, 2, false, true, false>(p, sp, pl, st);
    case 2: return launch_diag_t<2, 2, false, true, false>(p, sp, pl, st);
    case 3: return launch_diag_t<3, 2, false, true, false>(p, sp, pl, st);
    case 4: return launch_diag_t<4, 2, false, true, false>(p, sp, pl, st);
    case 5: return launch_diag_t<5, 2, false, true, false>(p, sp, pl, st);
    case 6: return launch_diag_t<6, 2, false, true, false>(p, sp, pl, st);
  }
  return SEQ_ERR_UNSUPPORTED;
}
#endif

}  // namespace seq
