// SEQ_METHOD_DIAG for kets ("ketdiag"): the sesolve path - psi' = -i (H0 + sum_k c_k(t) H_k) psi, reached from
// CompiledPulseSequence.run (basic.py:504) for ket inputs without collapse operators, which is what most of the
// reference's own tests and its qasm examples integrate (test_qasm.py:463-530: five qutrits, N = 243) - with the
// operators stored by diagonals, as kernel_diag.cuh does for density matrices:
//     (H psi)_i = w[i] psi_i + sum_t c_t(t) v_t[i] psi[i + o_t]
// one term per (operator, occupied off-diagonal); w = H0's main diagonal; c_t = 1 for H0's terms, the channel's
// spline value otherwise.  A drive a + a^dag of one mode is two terms, so a five-qutrit register costs ~21
// multiply-adds per element and right-hand side where the dense product costs N = 243 plus the N^2 (1 + n_ch)
// accumulation of G(t) (the generic kernel: 113 trajectories/s at N = 243; this kernel: see DESIGN.md).
// One CTA per trajectory, every vector (stage input with zero guard bands, y, k1..k7) in shared memory, element i
// on thread i mod blockDim; rolled stage loop; the shared step controller (StepCtl via diag_controller) on warp 0.
// sesolve semantics as in the generic kernel: with SEQ_FLAG_NORMALIZE the state is renormalised at every output and
// the integrator restarts from it (k1 is re-evaluated).
#pragma once
#include <cstdio>
#include <cstring>
#include "kernel_diag.cuh"

namespace seq {

#define KET_MAXT 96        // terms
#define KET_MAX_EPT 8

struct KetOps {
  int nTerm, nTab, has_w, maxoff;
  int t_tab[KET_MAXT], t_off[KET_MAXT], t_ch[KET_MAXT];
  int nnzE, stage_e;
  int tpc, nt;               // trajectories per CTA (they share the tables), threads per trajectory
  const cplx* tabs;          // [nTab][N]; row 0: H0's main diagonal
  const int* e_off;          // COO lists of the expectation operators (layout shared with the ELL / diagonal kernels)
  const unsigned* e_ij;
  const cplx* e_val;
};

// flags[op][o + N - 1] = 1 if operator op (0: H0, 1 + k: H_k) has a non-zero on diagonal o
__global__ void ket_flags_kernel(const cplx* __restrict__ H0, const cplx* __restrict__ Hk, int N, int* __restrict__ flags) {
  const int op = blockIdx.y;
  const long long N2 = (long long)N * N;
  const cplx* A = op == 0 ? H0 : Hk + (long long)(op - 1) * N2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N2; t += (long long)gridDim.x * blockDim.x) {
    const cplx v = A[t];
    if (v.x != 0.0 || v.y != 0.0) {
      const int i = (int)(t / N), j = (int)(t - (long long)i * N);
      flags[(long long)op * (2 * N - 1) + (j - i + N - 1)] = 1;
    }
  }
}

struct KetFill { int nTab; short op[KET_MAXT + 1]; int off[KET_MAXT + 1]; };

// tabs[t][i] = Op_t[i, i + off_t] (0 outside the matrix)
__global__ void ket_fill_kernel(const __grid_constant__ KetFill f, const cplx* __restrict__ H0, const cplx* __restrict__ Hk, int N, cplx* __restrict__ tabs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (i >= N) return;
  const long long N2 = (long long)N * N;
  const cplx* A = f.op[t] == 0 ? H0 : Hk + (long long)(f.op[t] - 1) * N2;
  const int j = i + f.off[t];
  tabs[(long long)t * N + i] = (j >= 0 && j < N) ? A[(long long)i * N + j] : cmake(0, 0);
}

struct KetLayout { int off_tabs, off_e, traj0, traj_bytes, off_Y, off_y, off_k, off_cv, off_red, off_ctrl, total; };      // off_Y.. are relative to a trajectory's region
__host__ __device__ inline KetLayout ket_layout(int N, int n_g, const KetOps& ko) {
  KetLayout L;
  const int c = (int)sizeof(cplx);
  // shared by the trajectories of a CTA: the diagonal tables and the expectation lists
  L.off_tabs = 0;
  L.off_e = ko.nTab * N * c;
  L.traj0 = L.off_e + (ko.stage_e ? (ko.nnzE > 0 ? ko.nnzE : 1) * 32 : 0);
  // per trajectory: stage input between zero guard bands of maxoff elements, y, k1..k7, spline values, reduction scratch, control
  L.off_Y = ko.maxoff * c;
  L.off_y = L.off_Y + N * c + ko.maxoff * c;
  L.off_k = L.off_y + N * c;
  L.off_cv = L.off_k + 7 * N * c;
  L.off_red = L.off_cv + 7 * (n_g > 0 ? n_g : 1) * (int)sizeof(double);
  L.off_red = (L.off_red + 15) / 16 * 16;
  L.off_ctrl = L.off_red + 64 * (int)sizeof(double);
  L.traj_bytes = (L.off_ctrl + 128 + 192 + 15) / 16 * 16;
  L.total = L.traj0 + ko.tpc * L.traj_bytes;
  return L;
}

// named barrier of one trajectory's threads (0 is __syncthreads)
__device__ __forceinline__ void ket_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// Up to two trajectories per CTA: the diagonal tables (the largest shared-memory item: terms x N x 16 B) are shared, every
// trajectory has its own vectors, controller and named barrier - twice the warps per SM for a latency-bound kernel.
template <int EPT>
__global__ void __launch_bounds__(512) solve_ketdiag_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ KetOps ko) {
  extern __shared__ __align__(16) unsigned char smem_ket[];
  const int N = p.N, n_g = p.n_g, nth = ko.nt;
  const int half = (int)threadIdx.x / nth, tid = (int)threadIdx.x - half * nth;
  const int b = blockIdx.x * ko.tpc + half;
  const bool live = b < p.B;                       // (an odd batch leaves the last CTA's second half idle)
  const int lane = tid & 31, warp = tid >> 5, nw = nth >> 5;
  const KetLayout L = ket_layout(N, n_g, ko);
  cplx* tabs = (cplx*)(smem_ket + L.off_tabs);
  unsigned char* mine = smem_ket + L.traj0 + half * L.traj_bytes;
  cplx* Ys = (cplx*)(mine + L.off_Y);
  cplx* y = (cplx*)(mine + L.off_y);
  cplx* kv = (cplx*)(mine + L.off_k);
  double* cv = (double*)(mine + L.off_cv);
  double* red = (double*)(mine + L.off_red);
  DiagCtrl* ctrl = (DiagCtrl*)(mine + L.off_ctrl);
  DiagCtlState* cst = (DiagCtlState*)(mine + L.off_ctrl + 128);
  const int bb = live ? b : 0;
  const double2* tabc = p.tab + (long long)bb * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)bb * p.tab_bstride_r;
  const double* cscale_b = p.coeff_scale ? p.coeff_scale + (long long)bb * p.n_ch : nullptr;
  auto sync_traj = [&]() { ket_bar(1 + half, nth); };

  for (int q = threadIdx.x; q < L.total / 16; q += blockDim.x) ((cplx*)smem_ket)[q] = cmake(0, 0);      // guard bands and everything else
  __syncthreads();
  for (int q = threadIdx.x; q < ko.nTab * N; q += blockDim.x) tabs[q] = ko.tabs[q];
  const unsigned* eij = ko.e_ij;
  const cplx* eval_ = ko.e_val;
  if (ko.stage_e) {
    cplx* ev_s = (cplx*)(smem_ket + L.off_e);
    unsigned* ei_s = (unsigned*)(ev_s + (ko.nnzE > 0 ? ko.nnzE : 1));
    for (int q = threadIdx.x; q < ko.nnzE; q += blockDim.x) { ev_s[q] = ko.e_val[q]; ei_s[q] = ko.e_ij[q]; }
    eij = ei_s; eval_ = ev_s;
  }
  __syncthreads();
  if (!live) return;                               // (after the CTA-wide barriers: the other half only uses its own named barrier from here on)
  const cplx* y0 = p.state0 + (long long)b * N;
  for (int i = tid; i < N; i += nth) { y[i] = y0[i]; Ys[i] = y0[i]; }
  if (tid == 0) {
    StepCtl c0;
    ctl_init(c0, p);
    cst->ctl = c0;
    cst->htry = 0.0; cst->target = 0.0; cst->o = 0; cst->nst = 0; cst->landing = 0; cst->capped = 0; cst->dbg = 0; cst->nstA = 0;
  }
  sync_traj();
  if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale_b, 0, 0, 0.0, 0);
  sync_traj();

  auto slot = [&](int j) -> cplx* { return kv + j * N; };      // slot j: k_{j+1}; FSAL: the last stage's slot is copied into slot 0 on commit
  // K = -i H(t) Ys for this thread's elements, stage values cvs[0..n_g)
  auto rhs = [&](const double* cvs, cplx* __restrict__ K) {
    cplx acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int i = tid + e * nth;
      acc[e] = (ko.has_w && i < N) ? cmul(tabs[i], Ys[i]) : cmake(0, 0);
    }
#pragma unroll 1
    for (int t = 0; t < ko.nTerm; t++) {
      const cplx* tv = tabs + ko.t_tab[t] * N;
      const int off = ko.t_off[t], ch = ko.t_ch[t];
      const double c = ch < 0 ? 1.0 : cvs[ch];
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int i = tid + e * nth;
        if (i < N) cfma(acc[e], seq::cscale(c, tv[i]), Ys[i + off]);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int i = tid + e * nth;
      if (i < N) K[i] = cmake(acc[e].y, -acc[e].x);
    }
  };
  auto emit = [&](int o) {
    if (p.flags & SEQ_FLAG_NORMALIZE) {
      double s = 0.0;
      for (int i = tid; i < N; i += nth) s += cabs2(y[i]);
      s = warp_sum(s);
      sync_traj();                 // (red may still be read by the error sum of the step before)
      if (lane == 0) red[32 + warp] = s;
      sync_traj();
      s = 0.0;
      for (int q = 0; q < nw; q++) s += red[32 + q];
      const double inv = 1.0 / sqrt(s);
      for (int i = tid; i < N; i += nth) y[i] = seq::cscale(inv, y[i]);
      sync_traj();
    }
    for (int e = warp; e < p.n_e; e += nw) {      // <psi|E|psi> = sum_ij conj(psi_i) E_ij psi_j, one warp per operator
      cplx acc = cmake(0, 0);
      const int q0 = ko.e_off[e], q1 = ko.e_off[e + 1];
      for (int q = q0 + lane; q < q1; q += 32) {
        const unsigned code = eij[q];                  // E_ij with i = code & 0xffff, j = code >> 16
        cplx row = cmul(eval_[q], y[code >> 16]);
        cfma_conj(acc, row, y[code & 0xffffu]);
      }
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (lane == 0) {
        const long long off = ((long long)b * p.n_e + e) * p.n_out + o;
        if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = acc;
        else ((double*)p.expect)[off] = acc.x;
      }
    }
    if (p.flags & SEQ_FLAG_STORE_STATES) {
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N;
      for (int i = tid; i < N; i += nth) dst[i] = y[i];
    }
  };

  int tog = 0, Slast = 7;
  bool restart = false;      // the state was renormalised at an output: k1 must be re-evaluated (sesolve restarts its integrator)
  while (true) {
    const DiagCtrl c = *ctrl;
    if (c.accepted) {        // commit: y <- new state, k1 <- derivative of the last stage
      for (int i = tid; i < N; i += nth) { y[i] = Ys[i]; kv[i] = kv[(Slast - 1) * N + i]; }
      sync_traj();
    }
    for (int o = c.emit_lo; o < c.emit_hi; o++) {
      emit(o);
      if (p.flags & SEQ_FLAG_NORMALIZE) restart = true;
    }
    if (c.cmd) break;
    const int pair = c.pair, S = pair ? 5 : 7;
    const int st0 = restart ? 0 : c.st0;
    restart = false;
    Slast = S;
    const double htry = c.htry;
#pragma unroll 1
    for (int st = st0; st < S; st++) {
      // stage input: Ys = y + h sum_{j<st} a_j k_j   (st == S - 1: the new state)
      for (int i = tid; i < N; i += nth) {
        double ax = 0.0, ay = 0.0;
        for (int j = 0; j < st; j++) {
          const double aj = RK_A[pair][st][j];
          const cplx kk = slot(j)[i];
          ax = fma(aj, kk.x, ax);
          ay = fma(aj, kk.y, ay);
        }
        const cplx v = y[i];
        Ys[i] = cmake(fma(htry, ax, v.x), fma(htry, ay, v.y));
      }
      sync_traj();
      rhs(cv + st * n_g, slot(st));
      sync_traj();
    }
    // error estimate (weighted RMS norm, as every other kernel)
    double es = 0.0;
    for (int i = tid; i < N; i += nth) {
      double ex = 0.0, ey = 0.0;
      for (int j = 0; j < S; j++) {
        const double ej = RK_E[pair][j];
        const cplx kk = slot(j)[i];
        ex = fma(ej, kk.x, ex);
        ey = fma(ej, kk.y, ey);
      }
      ex *= htry; ey *= htry;
      const double sc = p.atol + p.rtol * sqrt(fmax(cabs2(y[i]), cabs2(Ys[i])));
      es += (ex * ex + ey * ey) / (sc * sc);
    }
    es = warp_sum(es);
    if (lane == 0) red[tog * 32 + warp] = es;
    sync_traj();
    es = 0.0;
    for (int q = 0; q < nw; q++) es += red[tog * 32 + q];
    tog ^= 1;
    // (diag_controller normalises the error sum by N^2, the element count of a density matrix: a ket has N)
    if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale_b, 1, S - st0, es * (double)N, 0);
    sync_traj();
  }

  cplx* fin = p.final_state + (long long)b * N;
  for (int i = tid; i < N; i += nth) fin[i] = y[i];
  if (tid == 0) {
    p.status[b] = cst->ctl.status;
    p.stats[b * 4 + 0] = cst->ctl.nacc;
    p.stats[b * 4 + 1] = cst->ctl.nrej;
    p.stats[b * 4 + 2] = cst->ctl.nrhs;
    p.stats[b * 4 + 3] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct KetPlan {
  KetOps ops;
  KetFill fill;
  int threads, ept;
  size_t smem;
  long long macs;      // complex multiply-adds per element and right-hand side
};

// flags: ket_flags_kernel's output, [1 + n_ch][2N - 1]
static inline bool ket_plan(int N, int n_ch, int n_g, int nnzE, const int* flags, int max_smem, KetPlan* pl) {
  const int W = 2 * N - 1;
  if (N < 2 || N > 0x7fff) return false;
  KetOps& o = pl->ops;
  KetFill& f = pl->fill;
  memset(&o, 0, sizeof(o));
  memset(&f, 0, sizeof(f));
  // table 0: H0's main diagonal (always present, zeros if H0 has none)
  f.op[0] = 0; f.off[0] = 0;
  int nTab = 1;
  o.has_w = flags[N - 1] ? 1 : 0;
  o.nTerm = 0;
  int maxoff = 0;
  for (int op = 0; op <= n_ch; op++)
    for (int q = 0; q < W; q++) {
      if (!flags[(long long)op * W + q]) continue;
      const int off = q - (N - 1);
      if (op == 0 && off == 0) continue;
      if (o.nTerm >= KET_MAXT) return false;
      f.op[nTab] = (short)op; f.off[nTab] = off;
      o.t_tab[o.nTerm] = nTab; o.t_off[o.nTerm] = off; o.t_ch[o.nTerm] = op == 0 ? -1 : op - 1;
      o.nTerm++; nTab++;
      maxoff = abs(off) > maxoff ? abs(off) : maxoff;
    }
  o.nTab = nTab; f.nTab = nTab;
  o.maxoff = maxoff;
  o.nnzE = nnzE;
  pl->macs = o.nTerm + o.has_w;
  pl->threads = (N + 31) / 32 * 32;
  if (pl->threads > 256) pl->threads = 256;
  pl->ept = (N + pl->threads - 1) / pl->threads;
  if (pl->ept > KET_MAX_EPT) return false;
  o.nt = pl->threads;
  o.tpc = 2;                   // two trajectories per CTA when their vectors fit next to the shared tables
  o.stage_e = (nnzE > 0 && nnzE <= 2048) ? 1 : 0;
  if (ket_layout(N, n_g, o).total > max_smem) o.stage_e = 0;
  if (ket_layout(N, n_g, o).total > max_smem) { o.tpc = 1; o.stage_e = (nnzE > 0 && nnzE <= 2048) ? 1 : 0; }
  if (ket_layout(N, n_g, o).total > max_smem) o.stage_e = 0;
  if (ket_layout(N, n_g, o).total > max_smem) return false;
  pl->smem = (size_t)ket_layout(N, n_g, o).total;
  return true;
}

#ifndef SEQ_DIAG_NO_LAUNCH
template <int EPT>
static int launch_ket_t(const SolveParams& p, const KetPlan& pl, cudaStream_t st) {
  auto kern = solve_ketdiag_kernel<EPT>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return SEQ_ERR_CUDA;
  kern<<<(unsigned)((p.B + pl.ops.tpc - 1) / pl.ops.tpc), (unsigned)(pl.threads * pl.ops.tpc), pl.smem, st>>>(p, pl.ops);
  return cudaGetLastError() == cudaSuccess ? SEQ_OK : SEQ_ERR_CUDA;
}
static inline int launch_ket_kernel(const SolveParams& p, const KetPlan& pl, cudaStream_t st) {
  switch (pl.ept) {
    case 1: return launch_ket_t<1>(p, pl, st);
    case 2: return launch_ket_t<2>(p, pl, st);
    case 3: return launch_ket_t<3>(p, pl, st);
    case 4: return launch_ket_t<4>(p, pl, st);
    case 5: return launch_ket_t<5>(p, pl, st);
    case 6: return launch_ket_t<6>(p, pl, st);
    case 7: return launch_ket_t<7>(p, pl, st);
    case 8: return launch_ket_t<8>(p, pl, st);
  }
  return SEQ_ERR_UNSUPPORTED;
}
#endif

}  // namespace seq
