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
// address arithmetic per term (the host precomputes every offset in bytes), and the terms with offset zero
// (detunings, Kerr terms, dephasing) use the thread's own registers.  A coefficient is zero wherever its
// source would fall outside the matrix, so an out-of-range address only has to be readable and finite: Y
// sits between zero guard bands when they fit in shared memory (GUARD), otherwise the address is clamped.
//
// Work decomposition: one CTA per trajectory, each thread owns EPT elements of the upper triangle (y, RK
// stages, accumulators in registers), the full stage matrix Y double-buffered in shared memory, one barrier
// per right-hand side.  Step control (the shared StepCtl: identical decisions to every other kernel), the
// spline evaluation of all stage times of the next step and the output bookkeeping run on warp 0 only and
// reach the other warps through a small control block in shared memory - the scalar pow()/ceil()/division
// chains would otherwise be re-executed by every warp.  Operators that are row-sparse but not
// diagonal-structured take the ELL kernel (SEQ_METHOD_SPARSE), dense ones the tensor-core kernels.
#pragma once
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
  // --- collapse pair terms on the own element / on a gathered element
  int nLo, nLg;
  int lo_a[DIAG_MAX_LT], lo_b[DIAG_MAX_LT], lo_r[DIAG_MAX_LT];   // table a, table b (bytes inside lam); rate index or -1
  int lg_a[DIAG_MAX_LT], lg_b[DIAG_MAX_LT], lg_r[DIAG_MAX_LT], lg_off[DIAG_MAX_LT];
  int nT;                     // collapse diagonals (tables)
  int n_el, ld, ybufs, guard; // owned elements, leading dimension, Y buffers, guard band (elements)
  int stage_g, stage_e, stage_tab, nnzE;   // what is staged in shared memory
  const cplx* g0;             // [nD][N]
  const cplx* gk;             // [n_g][nD][N]
  const cplx* lam;            // [nT][N]
  const unsigned* ij;         // [n_el]
  const int* e_off;           // COO lists of the expectation operators (layout shared with the ELL kernel)
  const unsigned* e_ij;
  const cplx* e_val;
};

// ---------------------------------------------------------------------------------------------------
// preparation
// ---------------------------------------------------------------------------------------------------
// flags[op][o + N - 1] = 1 if operator `op` (0: union of G0, Gk; 1..n_l: L_l) has a non-zero on diagonal o
__global__ void diag_flags_kernel(const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                  const cplx* __restrict__ L, int N, int* __restrict__ flags) {
  const int op = blockIdx.y;
  const long long N2 = (long long)N * N;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N2; t += (long long)gridDim.x * blockDim.x) {
    bool nz;
    if (op == 0) {
      cplx v = G0[t];
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
                                 cplx* __restrict__ lam, int* __restrict__ any_imag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
  if (i >= N) return;
  const long long N2 = (long long)N * N;
  if (d < f.nD) {
    const int j = i + f.goff[d];
    const bool in = j >= 0 && j < N;
    g0[(long long)d * N + i] = in ? G0[(long long)i * N + j] : cmake(0, 0);
    for (int m = 0; m < n_g; m++) gk[((long long)m * f.nD + d) * N + i] = in ? Gk[m * N2 + (long long)i * N + j] : cmake(0, 0);
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
};

// controller state, kept in shared memory between calls (warp 0 only)
struct DiagCtlState {
  StepCtl ctl;
  double htry, target;
  int o, nst, landing, capped;
};

struct DiagLayout {
  int off_Y, ystride, ybytes, off_Gt, gtbytes, off_lam, off_g, off_e, off_tab, off_cv, off_red, off_ctrl, total;
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
  L.off_red = L.off_cv + 7 * (n_g > 0 ? n_g : 1) * (int)sizeof(double);
  L.off_ctrl = L.off_red + 64 * (int)sizeof(double);
  L.total = L.off_ctrl + 64 + 192;   // DiagCtrl + DiagCtlState
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
                                             int update, int nrhs, double es) {
  const int lane = threadIdx.x & 31, n_g = p.n_g;
  StepCtl ctl = S->ctl;
  double c_htry = S->htry, c_target = S->target;
  int c_o = S->o, c_nst = S->nst;
  bool c_landing = S->landing != 0, c_capped = S->capped != 0;
  __syncwarp();
  bool accepted = false;
  if (update) {
    ctl.nrhs += nrhs;
    ctl.have_k1 = true;
    const double err = sqrt(es / ((double)p.N * (double)p.N));
    accepted = ctl_update(ctl, p, err, c_htry, c_landing, c_capped, c_target, c_nst);
  }
  const int emit_lo = c_o;
  int cmd = 1;
  while (ctl.status == SEQ_STATUS_OK && c_o < p.n_out) {
    c_target = p.t0 + p.dt * (double)p.out_idx[c_o];
    if (c_o > 0 && ctl_next(ctl, p, c_target, c_htry, c_landing, c_capped)) { cmd = 0; break; }
    c_o++;           // output c_o is reached by the committed state
    c_nst = 0;
  }
  const int emit_hi = (ctl.status == SEQ_STATUS_OK) ? c_o : emit_lo;
  const int pair = ctl.pair;
  if (cmd == 0) {
    for (int q = lane; q < 7 * n_g; q += 32) {
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
  if (lane == 0) {
    ctrl->htry = c_htry;
    ctrl->cmd = cmd;
    ctrl->pair = pair;
    ctrl->st0 = ctl.have_k1 ? 1 : 0;
    ctrl->accepted = accepted ? 1 : 0;
    ctrl->emit_lo = emit_lo;
    ctrl->emit_hi = emit_hi;
    S->ctl = ctl;
    S->htry = c_htry; S->target = c_target;
    S->o = c_o; S->nst = c_nst; S->landing = c_landing ? 1 : 0; S->capped = c_capped ? 1 : 0;
  }
  __syncwarp();
}

#define DIAG_MAXT(EPT, KREG) ((KREG) ? ((EPT) <= 2 ? 544 : ((EPT) == 3 ? 384 : 512)) : ((EPT) <= 4 ? 1024 : ((EPT) <= 6 ? 704 : 512)))

template <int EPT, bool KREG, bool GUARD, bool LREAL>
__global__ void __launch_bounds__(DIAG_MAXT(EPT, KREG), 1)
solve_diag_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ DiagOps sp) {
  extern __shared__ __align__(16) unsigned char smem_diag[];
  const int N = p.N, ld = sp.ld, n_g = p.n_g;
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const DiagLayout L = diag_layout(N, n_g, p.n_t, sp);
  double* cv = (double*)(smem_diag + L.off_cv);
  double* red = (double*)(smem_diag + L.off_red);
  DiagCtrl* ctrl = (DiagCtrl*)(smem_diag + L.off_ctrl);

  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;
  cplx* kst = KREG ? nullptr : p.ws + (long long)b * p.ws_stride;   // [6][EPT][nth]

  // ---- stage shared tables: zero guard bands / Y padding, collapse diagonals, G diagonals, E, spline tables
  for (int q = tid; q < L.off_Gt / 16; q += nth) ((cplx*)smem_diag)[q] = cmake(0, 0);
  for (int q = tid; q < sp.nT * N; q += nth) ((cplx*)(smem_diag + L.off_lam))[q] = sp.lam[q];
  const int gsz = sp.nD * N;
  const cplx* g0p = sp.g0;
  const cplx* gkp = sp.gk;
  if (sp.stage_g) {
    cplx* gs = (cplx*)(smem_diag + L.off_g);
    for (int q = tid; q < gsz; q += nth) gs[q] = sp.g0[q];
    for (int q = tid; q < n_g * gsz; q += nth) gs[gsz + q] = sp.gk[q];
    g0p = gs; gkp = gs + gsz;
  }
  const unsigned* eij = sp.e_ij;
  const cplx* eval_ = sp.e_val;
  if (sp.stage_e) {
    cplx* ev_s = (cplx*)(smem_diag + L.off_e);
    unsigned* ei_s = (unsigned*)(ev_s + (sp.nnzE > 0 ? sp.nnzE : 1));
    for (int q = tid; q < sp.nnzE; q += nth) { ev_s[q] = sp.e_val[q]; ei_s[q] = sp.e_ij[q]; }
    eij = ei_s; eval_ = ev_s;
  }
  if (sp.stage_tab) {
    double2* ts = (double2*)(smem_diag + L.off_tab);
    for (int q = tid; q < p.n_ch * p.n_t; q += nth) ts[q] = tabc[q];
    for (int q = tid; q < p.n_ctd * p.n_t; q += nth) ts[p.n_ch * p.n_t + q] = tabr[q];
    tabc = ts; tabr = ts + p.n_ch * p.n_t;
  }

  // ---- owned elements: byte offsets of Y_ij inside a buffer, of row-table entry i and of entry j.  The
  // offsets are laundered through an empty asm so that they stay in registers as they are instead of being
  // re-derived from (i, j) at every use.
  int oy[EPT], oi[EPT], oj[EPT];
  unsigned vmask = 0;
  cplx y[EPT], ys[EPT], kout[EPT];
  cplx k[KREG ? 6 : 1][EPT];
  const cplx* y0 = p.state0 + (long long)b * N * N;
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    const int idx = tid + e * nth;
    const bool v = idx < sp.n_el;
    const unsigned code = v ? sp.ij[idx] : 0u;
    const int i = code & 0xffffu, j = code >> 16;
    if (v) vmask |= 1u << e;
    oy[e] = (i * ld + j) * (int)sizeof(cplx);
    oi[e] = i * (int)sizeof(cplx);
    oj[e] = j * (int)sizeof(cplx);
    y[e] = v ? y0[i * N + j] : cmake(0, 0);
    asm volatile("" : "+r"(oy[e]), "+r"(oi[e]), "+r"(oj[e]));
    ys[e] = y[e];
    kout[e] = cmake(0, 0);
#pragma unroll
    for (int q = 0; q < (KREG ? 6 : 1); q++) k[q][e] = cmake(0, 0);
  }
  __syncthreads();   // guard bands zeroed before anything is written into Y

  int yb = 0, gb = 0, tog = 0, ycur = 0;
  auto write_Y = [&](unsigned char* Yb, const cplx (&v)[EPT]) {
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      if (vmask & (1u << e)) {
        *(cplx*)(Yb + oy[e]) = v[e];
        if (oi[e] != oj[e]) *(cplx*)(Yb + oj[e] * ld + oi[e]) = cconj(v[e]);
      }
    }
  };
  auto build_Gt = [&](cplx* Gb, int st) {
    for (int q = tid; q < gsz; q += nth) {
      cplx g = g0p[q];
      for (int m = 0; m < n_g; m++) {
        const cplx a = gkp[m * gsz + q];
        const double v = cv[st * n_g + m];
        g.x = fma(v, a.x, g.x);
        g.y = fma(v, a.y, g.y);
      }
      Gb[q] = g;
    }
  };
  const int y_hi = L.ybytes - (int)sizeof(cplx);
  auto yaddr = [&](int a) -> int { return GUARD ? a : min(max(a, 0), y_hi); };

  auto compute_K = [&](const unsigned char* __restrict__ Yb, const unsigned char* __restrict__ Gb, int st) {
    cplx D[EPT];
    if (sp.g0_tab >= 0) {
      const unsigned char* gv = Gb + sp.g0_tab;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const cplx vi = *(const cplx*)(gv + oi[e]);
        const cplx vj = *(const cplx*)(gv + oj[e]);
        const cplx dd = cmake(vi.x - vj.x, vi.y + vj.y);   // v_0[i] - conj(v_0[j])
        D[e] = cmul(dd, ys[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < EPT; e++) D[e] = cmake(0, 0);
    }
#pragma unroll 1
    for (int t = 0; t < sp.nGo; t++) {
      const unsigned char* gv = Gb + sp.go_tab[t];
      const unsigned char* YL = Yb + sp.go_offL[t];
      const unsigned char* YR = Yb + sp.go_offR[t];
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const cplx vi = *(const cplx*)(gv + oi[e]);
        const cplx ya = GUARD ? *(const cplx*)(YL + oy[e]) : *(const cplx*)(Yb + yaddr(oy[e] + sp.go_offL[t]));
        cfma(D[e], vi, ya);
        const cplx vj = *(const cplx*)(gv + oj[e]);
        const cplx yv = GUARD ? *(const cplx*)(YR + oy[e]) : *(const cplx*)(Yb + yaddr(oy[e] + sp.go_offR[t]));
        // D -= conj(vj) * yv
        D[e].x = fma(-vj.x, yv.x, D[e].x); D[e].x = fma(-vj.y, yv.y, D[e].x);
        D[e].y = fma(-vj.x, yv.y, D[e].y); D[e].y = fma(vj.y, yv.x, D[e].y);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; e++) kout[e] = cmake(D[e].y, -D[e].x);
    const unsigned char* lam = smem_diag + L.off_lam;
    // collapse terms on the own element (dephasing-like: both diagonals are the main diagonal)
#pragma unroll 1
    for (int t = 0; t < sp.nLo; t++) {
      const unsigned char* la = lam + sp.lo_a[t];
      const unsigned char* lb = lam + sp.lo_b[t];
      const int ri = sp.lo_r[t];
      const double r = ri < 0 ? 1.0 : cv[st * n_g + ri];
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        if (LREAL) {
          const double w = r * (*(const double*)(la + oi[e])) * (*(const double*)(lb + oj[e]));
          kout[e].x = fma(w, ys[e].x, kout[e].x);
          kout[e].y = fma(w, ys[e].y, kout[e].y);
        } else {
          const cplx a = *(const cplx*)(la + oi[e]), c = *(const cplx*)(lb + oj[e]);
          const double wx = r * (a.x * c.x + a.y * c.y), wy = r * (a.y * c.x - a.x * c.y);
          kout[e].x = fma(wx, ys[e].x, kout[e].x); kout[e].x = fma(-wy, ys[e].y, kout[e].x);
          kout[e].y = fma(wx, ys[e].y, kout[e].y); kout[e].y = fma(wy, ys[e].x, kout[e].y);
        }
      }
    }
    // collapse terms that gather rho[i + o_a, j + o_b]
#pragma unroll 1
    for (int t = 0; t < sp.nLg; t++) {
      const unsigned char* la = lam + sp.lg_a[t];
      const unsigned char* lb = lam + sp.lg_b[t];
      const unsigned char* YG = Yb + sp.lg_off[t];
      const int ri = sp.lg_r[t];
      const double r = ri < 0 ? 1.0 : cv[st * n_g + ri];
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const cplx yv = GUARD ? *(const cplx*)(YG + oy[e]) : *(const cplx*)(Yb + yaddr(oy[e] + sp.lg_off[t]));
        if (LREAL) {
          const double w = r * (*(const double*)(la + oi[e])) * (*(const double*)(lb + oj[e]));
          kout[e].x = fma(w, yv.x, kout[e].x);
          kout[e].y = fma(w, yv.y, kout[e].y);
        } else {
          const cplx a = *(const cplx*)(la + oi[e]), c = *(const cplx*)(lb + oj[e]);
          const double wx = r * (a.x * c.x + a.y * c.y), wy = r * (a.y * c.x - a.x * c.y);
          kout[e].x = fma(wx, yv.x, kout[e].x); kout[e].x = fma(-wy, yv.y, kout[e].x);
          kout[e].y = fma(wx, yv.y, kout[e].y); kout[e].y = fma(wy, yv.x, kout[e].y);
        }
      }
    }
  };
  auto emit = [&](const unsigned char* __restrict__ Yb, int o) {
    const int nw = (nth + 31) >> 5;
    for (int e = warp; e < p.n_e; e += nw) {
      cplx acc = cmake(0, 0);
      const int q1 = sp.e_off[e + 1];
      for (int q = sp.e_off[e] + lane; q < q1; q += 32) {
        const unsigned code = eij[q];
        cfma(acc, eval_[q], ((const cplx*)Yb)[(code >> 16) * ld + (code & 0xffffu)]);   // E_ij rho_ji
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
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N * N;
      for (int q = tid; q < N * N; q += nth) { const int i = q / N, j = q - i * N; dst[q] = ((const cplx*)Yb)[i * ld + j]; }
    }
  };

  DiagCtlState* cst = (DiagCtlState*)(smem_diag + L.off_ctrl + 64);
  if (tid == 0) {
    StepCtl c0;
    ctl_init(c0, p);
    cst->ctl = c0;
    cst->htry = 0.0; cst->target = 0.0; cst->o = 0; cst->nst = 0; cst->landing = 0; cst->capped = 0;
  }
  __syncwarp();

  write_Y(smem_diag + L.off_Y, y);
  if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 0, 0, 0.0);
  __syncthreads();

  // One embedded-RK step with the stage loop unrolled at compile time: static register indices for the stage
  // derivatives, tableau coefficients as immediates (zero entries vanish).  Returns this thread's share of
  // the weighted squared error.
  auto run_step = [&](auto tag, const int st0, const double htry) -> double {
    constexpr int PAIR = decltype(tag)::value;
    constexpr int S = PAIR ? 5 : 7;
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
            const cplx kv = KREG ? k[KREG ? j : 0][e] : kst[((long long)j * EPT + e) * nth + tid];
            ax = fma(aj, kv.x, ax);
            ay = fma(aj, kv.y, ay);
          }
        }
        ys[e] = (st == 0) ? y[e] : cmake(fma(htry, ax, y[e].x), fma(htry, ay, y[e].y));
      }
      if (sp.ybufs == 2) yb ^= 1;
      gb ^= 1;
      unsigned char* Yb = smem_diag + L.off_Y + yb * L.ystride;
      unsigned char* Gb = smem_diag + L.off_Gt + gb * L.gtbytes;
      write_Y(Yb, ys);
      build_Gt((cplx*)Gb, st);
      __syncthreads();
      compute_K(Yb, Gb, st);
      if (st < S - 1) {
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          if (KREG) k[KREG ? st : 0][e] = kout[e];
          else kst[((long long)st * EPT + e) * nth + tid] = kout[e];
        }
      }
      if (sp.ybufs == 1) __syncthreads();       // all gathers done before the next stage overwrites Y
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
          const cplx kv = KREG ? k[KREG ? j : 0][e] : kst[((long long)j * EPT + e) * nth + tid];
          ex = fma(cj, kv.x, ex);
          ey = fma(cj, kv.y, ey);
        }
      }
      ex *= htry; ey *= htry;
      const double sc = p.atol + p.rtol * sqrt(fmax(cabs2(y[e]), cabs2(ys[e])));
      const double wgt = (vmask & (1u << e)) ? ((oi[e] == oj[e]) ? 1.0 : 2.0) : 0.0;
      es += wgt * (ex * ex + ey * ey) / (sc * sc);
    }
    return es;
  };

  while (true) {
    const DiagCtrl c = *ctrl;
    if (c.accepted) {
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        y[e] = ys[e];
        if (KREG) k[0][e] = kout[e];
        else kst[((long long)e) * nth + tid] = kout[e];
      }
      ycur = yb;
    }
    for (int o = c.emit_lo; o < c.emit_hi; o++) emit(smem_diag + L.off_Y + ycur * L.ystride, o);
    if (c.cmd) break;
    if (sp.ybufs == 1 && c.emit_hi > c.emit_lo) __syncthreads();   // emit reads the buffer the next stage overwrites
    double es;
    if (c.pair) es = run_step(IntTag<1>(), c.st0, c.htry);
    else es = run_step(IntTag<0>(), c.st0, c.htry);
    es = diag_block_sum(es, red, tog);
    if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 1, (c.pair ? 5 : 7) - c.st0, es);
    __syncthreads();
  }

  cplx* fin = p.final_state + (long long)b * N * N;
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    if (vmask & (1u << e)) {
      const int i = oi[e] / (int)sizeof(cplx), j = oj[e] / (int)sizeof(cplx);
      fin[i * N + j] = y[e];
      if (i != j) fin[j * N + i] = cconj(y[e]);
    }
  }
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
struct DiagPlan {
  DiagOps ops;          // offsets / term lists filled in; pointers set at launch
  DiagFill fill;
  int maxoff;
  int ept, threads;
  bool kreg, guard;
  size_t smem;
  long long macs;       // complex multiply-adds per element (cost model)
};

// flags: host copy of diag_flags_kernel's output, [1 + n_l][2N - 1].  Returns false if the operators are not
// diagonal-structured enough (too many diagonals) or no launch shape fits.
static inline bool diag_plan(int N, int n_g, int n_c, int n_ctd, int n_ch, int n_t, int nnzE, const int* flags, int max_smem,
                             int ept_override, DiagPlan* pl) {
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
  if (o.nD == 0) { f.goff[0] = 0; o.g0_tab = 0; o.nD = 1; }   // G == 0: keep one (zero) diagonal
  f.nD = o.nD;
  int nT = 0;
  o.nLo = o.nLg = 0;
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
          if (o.nLo >= DIAG_MAX_LT) return false;
          o.lo_a[o.nLo] = a * N * c; o.lo_b[o.nLo] = b2 * N * c; o.lo_r[o.nLo] = rate; o.nLo++;
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
  pl->macs = 2LL * o.nD + o.nLo + o.nLg;
  o.n_el = N * (N + 1) / 2;
  // launch shape: registers hold k1..k6 for up to 4 elements per thread; beyond that the stage vectors
  // stream through a thread-private L2 buffer
  pl->ept = 0;
  const int kreg_epts[4] = {1, 2, 3, 4};
  for (int q = 0; q < 4 && !pl->ept; q++) {
    const int e = kreg_epts[q];
    if (ept_override > 0 && e != ept_override) continue;
    const long long th = (((long long)o.n_el + e - 1) / e + 31) / 32 * 32;
    const int cap = (e == 1) ? 128 : DIAG_MAXT(e, true);
    if (th <= cap || (ept_override > 0 && th <= DIAG_MAXT(e, true))) { pl->ept = e; pl->threads = (int)th; pl->kreg = true; }
  }
  if (!pl->ept) {
    const int st_epts[3] = {8, 6, 4};
    for (int q = 0; q < 3 && !pl->ept; q++) {
      const int e = st_epts[q];
      if (ept_override > 0 && e != ept_override) continue;
      const long long th = (((long long)o.n_el + e - 1) / e + 31) / 32 * 32;
      if (th <= DIAG_MAXT(e, false)) { pl->ept = e; pl->threads = (int)th; pl->kreg = false; }
    }
  }
  if (!pl->ept) return false;
  // shared memory, in order of value: Y (guard bands + double buffering, then clamped addresses, then a single
  // buffer), then the G diagonals, the spline tables and the expectation operators as far as they fit
  const int guard_el = maxoff * o.ld + maxoff + 1;
  const int tries[3][2] = {{2, 1}, {2, 0}, {1, 0}};
  bool ok = false;
  for (int q = 0; q < 3 && !ok; q++) {
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
  pl->smem = (size_t)diag_layout(N, n_g, n_t, o).total;
  return true;
}

static inline size_t diag_ws_elems(const DiagPlan& pl) { return pl.kreg ? 0 : (size_t)6 * pl.ept * pl.threads; }

struct DiagPack { size_t o_g0, o_gk, o_lam, o_flag, bytes; };
static inline DiagPack diag_pack_layout(int N, int n_g, const DiagPlan& pl) {
  auto up = [](size_t x) { return (x + 15) / 16 * 16; };
  DiagPack k;
  size_t o = 0;
  // g0 and gk contiguous: the kernel stages them with one layout
  k.o_g0 = o; o += (size_t)pl.ops.nD * N * sizeof(cplx);
  k.o_gk = o; o += up((size_t)(n_g > 0 ? n_g : 1) * pl.ops.nD * N * sizeof(cplx));
  k.o_lam = o; o += up((size_t)(pl.ops.nT > 0 ? pl.ops.nT : 1) * N * sizeof(cplx));
  k.o_flag = o; o += 16;
  k.bytes = o + 64;
  return k;
}

template <int EPT, bool KREG>
static int launch_diag_t(const SolveParams& p, const DiagOps& sp, const DiagPlan& pl, bool lreal, cudaStream_t st) {
#define DIAG_LAUNCH(G, R)                                                                                                         \
  do {                                                                                                                            \
    if (cudaFuncSetAttribute(solve_diag_kernel<EPT, KREG, G, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) \
      return SEQ_ERR_CUDA;                                                                                                        \
    solve_diag_kernel<EPT, KREG, G, R><<<p.B, pl.threads, pl.smem, st>>>(p, sp);                                                  \
  } while (0)
  if (pl.guard) { if (lreal) DIAG_LAUNCH(true, true); else DIAG_LAUNCH(true, false); }
  else { if (lreal) DIAG_LAUNCH(false, true); else DIAG_LAUNCH(false, false); }
#undef DIAG_LAUNCH
  return SEQ_OK;
}

static inline int launch_diag_kernel(const SolveParams& p, const DiagOps& sp, const DiagPlan& pl, bool lreal, cudaStream_t st) {
  if (pl.kreg) {
    switch (pl.ept) {
      case 1: return launch_diag_t<1, true>(p, sp, pl, lreal, st);
      case 2: return launch_diag_t<2, true>(p, sp, pl, lreal, st);
      case 3: return launch_diag_t<3, true>(p, sp, pl, lreal, st);
      case 4: return launch_diag_t<4, true>(p, sp, pl, lreal, st);
    }
  } else {
    switch (pl.ept) {
      case 4: return launch_diag_t<4, false>(p, sp, pl, lreal, st);
      case 6: return launch_diag_t<6, false>(p, sp, pl, lreal, st);
      case 8: return launch_diag_t<8, false>(p, sp, pl, lreal, st);
    }
  }
  return SEQ_ERR_UNSUPPORTED;
}

}  // namespace seq
