// SEQ_METHOD_SPARSE: density matrices whose operators are row-sparse (the reference's own case:
// System.H0 / Mode.a / Mode.n are Kronecker-embedded ladder and number operators with 1-3 non-zeros
// per row, modes.py:125-207, system.py:399-445; QuTiP integrates the same structure as a CSR SpMV).
//
// One CTA integrates one trajectory over the whole time span.  The operators are kept in ELL form
// (per row: `w` column indices + values, rows padded with zeros):
//     K_ij = -i sum_s G_is(t) Y[c_is, j]  +  i sum_s conj(G_js(t)) Y[i, c_js]
//            + sum_l r_l sum_{a,b} L_l[i,a] conj(L_l[j,b]) Y[c_ia, c_jb]
// i.e. (2 w_G + sum_l w_l^2) complex multiply-adds per element instead of N (1 + 2 n_l), on the FP64
// FMA pipe.  rho is Hermitian, so a thread owns EPT elements of the UPPER triangle: y, the RK stage
// derivatives and the accumulators of those elements live in its registers (KREG) or in a
// thread-private coalesced L2 stream (large N); the current stage matrix Y is kept in full in shared
// memory (the owner writes Y_ij and Y_ji = conj) so that every gather is a plain, contiguous read.
// One barrier per right-hand side (Y and G(t) are double-buffered), one per block reduction.
// The step controller is the shared StepCtl, so accept/reject decisions match the other kernels.
#pragma once
#include "common.cuh"

namespace seq {

#define SPARSE_MAX_WG 16
#define SPARSE_MAX_WL 4
#define SPARSE_MAXT_KREG 512     /* EPT = 4: 128 registers per thread */
#define SPARSE_MAXT_KREG3 384    /* EPT <= 3: 168 registers per thread */
#define SPARSE_MAXT_STREAM 800

struct SparseOps {
  int wG, wL;             // ELL widths: union pattern of G0 and every Gk; max over the collapse operators
  int n_el;               // N (N + 1) / 2 owned elements
  int ld;                 // leading dimension of Y in shared memory (odd: transposed stores conflict-free)
  int ybufs;              // 2: Y double-buffered, 1: single buffer (extra barrier per RHS)
  const int* gcol;        // [wG][N]
  const cplx* g0;         // [wG][N]
  const cplx* gk;         // [n_g][wG][N]
  const int* lcol;        // [n_l][wL][N]
  const cplx* lval;       // [n_l][wL][N]
  const unsigned* ij;     // [n_el] i | j << 16, row-major over the upper triangle
  const int* e_off;       // [n_e + 1] offsets into the COO lists of the expectation operators
  const unsigned* e_ij;   // i | j << 16
  const cplx* e_val;
};

// ---------------------------------------------------------------------------------------------------
// preparation: row statistics, ELL fill, COO fill
// ---------------------------------------------------------------------------------------------------
// op 0: union pattern of G0 and Gk[0..n_g); op 1..n_l: L_l; op 1+n_l..: E_e.
// grid (N, n_ops), 32 threads: rowcnt[op][i] = nnz of row i; maxw[op] = max over rows; total[op] = sum.
__global__ void sparse_count_kernel(const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                    const cplx* __restrict__ L, int n_l, const cplx* __restrict__ E, int N,
                                    int* __restrict__ rowcnt, int* __restrict__ maxw, int* __restrict__ total) {
  const int i = blockIdx.x, op = blockIdx.y, lane = threadIdx.x;
  const long long N2 = (long long)N * N;
  int cnt = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    int j = j0 + lane;
    bool nz = false;
    if (j < N) {
      long long off = (long long)i * N + j;
      if (op == 0) {
        cplx v = G0[off];
        nz = (v.x != 0.0 || v.y != 0.0);
        for (int m = 0; m < n_g && !nz; m++) { cplx w = Gk[m * N2 + off]; nz = (w.x != 0.0 || w.y != 0.0); }
      } else {
        const cplx* M = (op <= n_l) ? L + (long long)(op - 1) * N2 : E + (long long)(op - 1 - n_l) * N2;
        cplx v = M[off];
        nz = (v.x != 0.0 || v.y != 0.0);
      }
    }
    cnt += __popc(__ballot_sync(0xffffffffu, nz));
  }
  if (lane == 0) {
    rowcnt[(long long)op * N + i] = cnt;
    atomicMax(maxw + op, cnt);
    atomicAdd(total + op, cnt);
  }
}

// exclusive scan of the row counts of the expectation operators (one thread per operator; N <= a few thousand)
__global__ void sparse_escan_kernel(const int* __restrict__ rowcnt, int n_l, int n_e, int N, int* __restrict__ rowoff,
                                    int* __restrict__ e_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int e = 0; e < n_e; e++) {
      e_off[e] = run;
      const int* rc = rowcnt + (long long)(1 + n_l + e) * N;
      for (int i = 0; i < N; i++) { rowoff[(long long)e * N + i] = run; run += rc[i]; }
    }
    e_off[n_e] = run;
  }
}

// grid (N, n_ops), 32 threads: ordered compaction of row i of operator `op` into its ELL / COO slot.
__global__ void sparse_fill_kernel(const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                   const cplx* __restrict__ L, int n_l, const cplx* __restrict__ E, int N, int wG, int wL,
                                   int* __restrict__ gcol, cplx* __restrict__ g0, cplx* __restrict__ gk,
                                   int* __restrict__ lcol, cplx* __restrict__ lval, const int* __restrict__ rowoff,
                                   unsigned* __restrict__ e_ij, cplx* __restrict__ e_val, int op0) {
  const int i = blockIdx.x, op = blockIdx.y + op0, lane = threadIdx.x;
  const long long N2 = (long long)N * N;
  int base = 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    int j = j0 + lane;
    bool nz = false;
    long long off = (long long)i * N + j;
    if (j < N) {
      if (op == 0) {
        cplx v = G0[off];
        nz = (v.x != 0.0 || v.y != 0.0);
        for (int m = 0; m < n_g && !nz; m++) { cplx w = Gk[m * N2 + off]; nz = (w.x != 0.0 || w.y != 0.0); }
      } else {
        const cplx* M = (op <= n_l) ? L + (long long)(op - 1) * N2 : E + (long long)(op - 1 - n_l) * N2;
        cplx v = M[off];
        nz = (v.x != 0.0 || v.y != 0.0);
      }
    }
    unsigned bal = __ballot_sync(0xffffffffu, nz);
    int s = base + __popc(bal & ((1u << lane) - 1u));
    if (nz) {
      if (op == 0) {
        gcol[(long long)s * N + i] = j;
        g0[(long long)s * N + i] = G0[off];
        for (int m = 0; m < n_g; m++) gk[((long long)m * wG + s) * N + i] = Gk[m * N2 + off];
      } else if (op <= n_l) {
        lcol[((long long)(op - 1) * wL + s) * N + i] = j;
        lval[((long long)(op - 1) * wL + s) * N + i] = L[(long long)(op - 1) * N2 + off];
      } else {
        int e = op - 1 - n_l;
        int dst = rowoff[(long long)e * N + i] + s;
        e_ij[dst] = (unsigned)i | ((unsigned)j << 16);
        e_val[dst] = E[(long long)e * N2 + off];
      }
    }
    base += __popc(bal);
  }
  // pad: column = own row (always in range), value 0
  if (op == 0) {
    for (int s = base + lane; s < wG; s += 32) {
      gcol[(long long)s * N + i] = i;
      g0[(long long)s * N + i] = cmake(0, 0);
      for (int m = 0; m < n_g; m++) gk[((long long)m * wG + s) * N + i] = cmake(0, 0);
    }
  } else if (op <= n_l) {
    for (int s = base + lane; s < wL; s += 32) {
      lcol[((long long)(op - 1) * wL + s) * N + i] = i;
      lval[((long long)(op - 1) * wL + s) * N + i] = cmake(0, 0);
    }
  }
}

// idx -> (i, j), row-major over the upper triangle
__global__ void sparse_ij_kernel(unsigned* __restrict__ ij, int N) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * N) return;
  int i = t / N, j = t - i * N;
  if (j < i) return;
  int idx = i * N - (i * (i - 1)) / 2 + (j - i);
  ij[idx] = (unsigned)i | ((unsigned)j << 16);
}

// ---------------------------------------------------------------------------------------------------
// the integration kernel
// ---------------------------------------------------------------------------------------------------
struct SparseSmem {
  cplx* Y;        // [ybufs][N*ld]
  cplx* Gt;       // [2][wG*N]
  cplx* lval;     // [n_l*wL*N]
  double* cv;     // [7][n_g]
  double* red;    // [2][32]
  int* gcol;      // [wG*N]
  int* lcol;      // [n_l*wL*N]
};

static inline size_t sparse_smem_bytes(int N, int ld, int ybufs, int wG, int wL, int n_l, int n_g) {
  size_t b = 0;
  b += (size_t)ybufs * N * ld * sizeof(cplx);
  b += (size_t)2 * wG * N * sizeof(cplx);
  b += (size_t)n_l * wL * N * sizeof(cplx);
  b += (size_t)(7 * (n_g > 0 ? n_g : 1) + 64) * sizeof(double);
  b += (size_t)(wG * N + n_l * wL * N + 4) * sizeof(int);
  return b + 32;
}

__device__ __forceinline__ double block_sum_tog(double v, double* red, int& tog) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[tog * 32 + w] = v;
  __syncthreads();
  double s = 0.0;
  for (int q = 0; q < nw; q++) s += red[tog * 32 + q];
  tog ^= 1;
  return s;
}

template <int EPT, bool KREG, bool WL1>
__global__ void __launch_bounds__(KREG ? (EPT <= 3 ? SPARSE_MAXT_KREG3 : SPARSE_MAXT_KREG) : SPARSE_MAXT_STREAM, 1)
solve_sparse_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ SparseOps sp) {
  extern __shared__ __align__(16) unsigned char smem_sparse[];
  const int N = p.N, ld = sp.ld, wG = sp.wG, wL = sp.wL, n_g = p.n_g, n_l = p.n_c + p.n_ctd;
  const int tid = threadIdx.x, nth = blockDim.x, b = blockIdx.x;
  const int ysz = N * ld, gsz = wG * N;
  SparseSmem s;
  {
    unsigned char* q = smem_sparse;
    s.Y = (cplx*)q; q += (size_t)sp.ybufs * ysz * sizeof(cplx);
    s.Gt = (cplx*)q; q += (size_t)2 * gsz * sizeof(cplx);
    s.lval = (cplx*)q; q += (size_t)n_l * wL * N * sizeof(cplx);
    s.cv = (double*)q; q += (size_t)7 * (n_g > 0 ? n_g : 1) * sizeof(double);
    s.red = (double*)q; q += 64 * sizeof(double);
    s.gcol = (int*)q; q += (size_t)gsz * sizeof(int);
    s.lcol = (int*)q;
  }
  for (int i = tid; i < gsz; i += nth) s.gcol[i] = sp.gcol[i];
  for (int i = tid; i < n_l * wL * N; i += nth) { s.lcol[i] = sp.lcol[i]; s.lval[i] = sp.lval[i]; }

  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;
  cplx* kst = KREG ? nullptr : p.ws + (long long)b * p.ws_stride;   // [6][EPT][nth]

  // owned elements
  int ei[EPT], ej[EPT];
  bool ev[EPT];
  cplx y[EPT], ys[EPT], kout[EPT];
  cplx k[KREG ? 6 : 1][EPT];
  const cplx* y0 = p.state0 + (long long)b * N * N;
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    int idx = tid + e * nth;
    ev[e] = idx < sp.n_el;
    unsigned code = ev[e] ? sp.ij[idx] : 0u;
    ei[e] = code & 0xffffu;
    ej[e] = code >> 16;
    y[e] = ev[e] ? y0[ei[e] * N + ej[e]] : cmake(0, 0);
    ys[e] = y[e];
    kout[e] = cmake(0, 0);
#pragma unroll
    for (int q = 0; q < (KREG ? 6 : 1); q++) k[q][e] = cmake(0, 0);
  }

  int yb = 0, gb = 0, tog = 0;     // buffer toggles (CTA-uniform)
  auto write_Y = [&](cplx* Yb, const cplx (&v)[EPT]) {
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      if (ev[e]) {
        Yb[ei[e] * ld + ej[e]] = v[e];
        if (ei[e] != ej[e]) Yb[ej[e] * ld + ei[e]] = cconj(v[e]);
      }
    }
  };
  // G(t) values of stage `st` into Gt[gb]
  auto build_Gt = [&](cplx* Gb, int st) {
    for (int q = tid; q < gsz; q += nth) {
      cplx g = sp.g0[q];
      for (int m = 0; m < n_g; m++) {
        cplx a = sp.gk[(long long)m * gsz + q];
        double v = s.cv[st * n_g + m];
        g.x = fma(v, a.x, g.x);
        g.y = fma(v, a.y, g.y);
      }
      Gb[q] = g;
    }
  };
  auto compute_K = [&](const cplx* __restrict__ Yb, const cplx* __restrict__ Gb, int st) {
    cplx d[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) d[e] = cmake(0, 0);
#pragma unroll 2
    for (int q = 0; q < wG; q++) {
      const int* gc = s.gcol + q * N;
      const cplx* gv = Gb + q * N;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int i = ei[e], j = ej[e];
        cplx gi = gv[i];
        cplx ya = Yb[gc[i] * ld + j];
        cfma(d[e], gi, ya);
        cplx gj = gv[j];
        cplx yv = Yb[i * ld + gc[j]];
        // d -= conj(gj) * yv
        d[e].x = fma(-gj.x, yv.x, d[e].x); d[e].x = fma(-gj.y, yv.y, d[e].x);
        d[e].y = fma(-gj.x, yv.y, d[e].y); d[e].y = fma(gj.y, yv.x, d[e].y);
      }
    }
#pragma unroll
    for (int e = 0; e < EPT; e++) kout[e] = cmake(d[e].y, -d[e].x);
#pragma unroll 1
    for (int l = 0; l < n_l; l++) {
      const double r = (l < p.n_c) ? 1.0 : s.cv[st * n_g + p.n_ch + (l - p.n_c)];
      if (WL1) {
        const int* lc = s.lcol + l * N;
        const cplx* lv = s.lval + l * N;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          const int i = ei[e], j = ej[e];
          cplx la = lv[i], lb = lv[j];
          cplx yv = Yb[lc[i] * ld + lc[j]];
          // w = r * la * conj(lb)
          double wx = r * (la.x * lb.x + la.y * lb.y), wy = r * (la.y * lb.x - la.x * lb.y);
          kout[e].x = fma(wx, yv.x, kout[e].x); kout[e].x = fma(-wy, yv.y, kout[e].x);
          kout[e].y = fma(wx, yv.y, kout[e].y); kout[e].y = fma(wy, yv.x, kout[e].y);
        }
      } else {
        for (int a = 0; a < wL; a++) {
          const int* lca = s.lcol + (l * wL + a) * N;
          const cplx* lva = s.lval + (l * wL + a) * N;
          for (int c = 0; c < wL; c++) {
            const int* lcb = s.lcol + (l * wL + c) * N;
            const cplx* lvb = s.lval + (l * wL + c) * N;
#pragma unroll
            for (int e = 0; e < EPT; e++) {
              const int i = ei[e], j = ej[e];
              cplx la = lva[i], lb = lvb[j];
              cplx yv = Yb[lca[i] * ld + lcb[j]];
              double wx = r * (la.x * lb.x + la.y * lb.y), wy = r * (la.y * lb.x - la.x * lb.y);
              kout[e].x = fma(wx, yv.x, kout[e].x); kout[e].x = fma(-wy, yv.y, kout[e].x);
              kout[e].y = fma(wx, yv.y, kout[e].y); kout[e].y = fma(wy, yv.x, kout[e].y);
            }
          }
        }
      }
    }
  };
  // expectation values (one warp per operator, lanes over its COO list) + optional state store
  auto emit = [&](const cplx* __restrict__ Yb, int o) {
    const int lane = tid & 31, w = tid >> 5, nw = (nth + 31) >> 5;
    for (int e = w; e < p.n_e; e += nw) {
      cplx acc = cmake(0, 0);
      const int q1 = sp.e_off[e + 1];
      for (int q = sp.e_off[e] + lane; q < q1; q += 32) {
        unsigned code = sp.e_ij[q];
        cfma(acc, sp.e_val[q], Yb[(code >> 16) * ld + (code & 0xffffu)]);   // E_ij rho_ji
      }
      acc.x = warp_sum(acc.x);
      acc.y = warp_sum(acc.y);
      if (lane == 0) {
        long long off = ((long long)b * p.n_e + e) * p.n_out + o;
        if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = acc;
        else ((double*)p.expect)[off] = acc.x;
      }
    }
    if (p.flags & SEQ_FLAG_STORE_STATES) {
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N * N;
      for (int q = tid; q < N * N; q += nth) { int i = q / N, j = q - i * N; dst[q] = Yb[i * ld + j]; }
    }
  };

  write_Y(s.Y, y);
  __syncthreads();
  int ycur = 0;                    // buffer that holds y

  StepCtl ctl;
  ctl_init(ctl, p);
  for (int o = 0; o < p.n_out && ctl.status == SEQ_STATUS_OK; o++) {
    if (o > 0) {
      const double t_target = p.t0 + p.dt * (double)p.out_idx[o];
      int nst = 0;
      double htry;
      bool landing, capped;
      while (ctl.status == SEQ_STATUS_OK && ctl_next(ctl, p, t_target, htry, landing, capped)) {
        const int pair = ctl.pair, S = RK_STAGES[pair];
        // spline values of every stage time of this step
        for (int q = tid; q < 7 * n_g; q += nth) {
          const int st = q / n_g, m = q - st * n_g;
          const double ts = ctl.t + RK_C[pair][st] * htry;
          double v;
          if (m < p.n_ch) {
            v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, ts);
            if (cscale) v *= cscale[m];
          } else {
            v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, ts);
          }
          s.cv[q] = v;
        }
        __syncthreads();
        for (int st = ctl.have_k1 ? 1 : 0; st < S; st++) {
          // stage input ys = y + h sum_{j<st} a_j k_j   (st == S-1: ys = y_new)
#pragma unroll
          for (int e = 0; e < EPT; e++) {
            double ax = 0.0, ay = 0.0;
            if (KREG) {
#pragma unroll
              for (int j = 0; j < 6; j++) {
                const double aj = RK_A[pair][st][j];
                ax = fma(aj, k[j][e].x, ax);
                ay = fma(aj, k[j][e].y, ay);
              }
            } else {
              for (int j = 0; j < st; j++) {
                const double aj = RK_A[pair][st][j];
                cplx kv = kst[((long long)j * EPT + e) * nth + tid];
                ax = fma(aj, kv.x, ax);
                ay = fma(aj, kv.y, ay);
              }
            }
            ys[e] = cmake(fma(htry, ax, y[e].x), fma(htry, ay, y[e].y));
          }
          if (sp.ybufs == 2) yb ^= 1;
          gb ^= 1;
          cplx* Yb = s.Y + (size_t)yb * ysz;
          cplx* Gb = s.Gt + (size_t)gb * gsz;
          write_Y(Yb, ys);
          build_Gt(Gb, st);
          __syncthreads();
          compute_K(Yb, Gb, st);
          ctl.nrhs++;
          if (st < S - 1) {
            if (KREG) {
#pragma unroll
              for (int q = 0; q < 6; q++)
                if (q == st) {
#pragma unroll
                  for (int e = 0; e < EPT; e++) k[q][e] = kout[e];
                }
            } else {
#pragma unroll
              for (int e = 0; e < EPT; e++) kst[((long long)st * EPT + e) * nth + tid] = kout[e];
            }
          }
          if (sp.ybufs == 1) __syncthreads();   // all gathers done before the next stage overwrites Y
        }
        ctl.have_k1 = true;
        // error estimate: h (sum_{j<S-1} e_j k_j + e_{S-1} k_S), weighted RMS over all N^2 elements
        double es = 0.0;
        const double elast = RK_E[pair][S - 1];
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          double ex = elast * kout[e].x, ey = elast * kout[e].y;
          if (KREG) {
#pragma unroll
            for (int j = 0; j < 6; j++) {
              const double cj = (j < S - 1) ? RK_E[pair][j] : 0.0;
              ex = fma(cj, k[j][e].x, ex);
              ey = fma(cj, k[j][e].y, ey);
            }
          } else {
            for (int j = 0; j < S - 1; j++) {
              const double cj = RK_E[pair][j];
              cplx kv = kst[((long long)j * EPT + e) * nth + tid];
              ex = fma(cj, kv.x, ex);
              ey = fma(cj, kv.y, ey);
            }
          }
          ex *= htry; ey *= htry;
          const double sc = p.atol + p.rtol * sqrt(fmax(cabs2(y[e]), cabs2(ys[e])));
          const double wgt = ev[e] ? ((ei[e] == ej[e]) ? 1.0 : 2.0) : 0.0;
          es += wgt * (ex * ex + ey * ey) / (sc * sc);
        }
        es = block_sum_tog(es, s.red, tog);
        const double err = sqrt(es / ((double)N * (double)N));
        if (ctl_update(ctl, p, err, htry, landing, capped, t_target, nst)) {
#pragma unroll
          for (int e = 0; e < EPT; e++) y[e] = ys[e];
          if (KREG) {
#pragma unroll
            for (int e = 0; e < EPT; e++) k[0][e] = kout[e];
          } else {
#pragma unroll
            for (int e = 0; e < EPT; e++) kst[((long long)e) * nth + tid] = kout[e];
          }
          ycur = yb;
        }
      }
      if (ctl.status != SEQ_STATUS_OK) break;
    }
    emit(s.Y + (size_t)ycur * ysz, o);
  }
  // final state from the registers (valid whatever the buffers hold after a failed step)
  cplx* fin = p.final_state + (long long)b * N * N;
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    if (ev[e]) {
      fin[ei[e] * N + ej[e]] = y[e];
      if (ei[e] != ej[e]) fin[ej[e] * N + ei[e]] = cconj(y[e]);
    }
  }
  if (tid == 0) {
    p.status[b] = ctl.status;
    p.stats[b * 4 + 0] = ctl.nacc;
    p.stats[b * 4 + 1] = ctl.nrej;
    p.stats[b * 4 + 2] = ctl.nrhs;
    p.stats[b * 4 + 3] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------
// host side: plan (widths -> layout), pack-buffer carving, launch
// ---------------------------------------------------------------------------------------------------
struct SparsePlan {
  int wG, wL, nnzE;      // from the device analysis
  long long macs;        // complex multiply-adds per element (cost model)
  int ept, threads;      // launch shape
  bool kreg;
  int ld, ybufs;
  size_t smem;
};

// sparse beats the dense tensor-core kernels when the per-element multiply-add count is far below N (1 + 2 n_l)
static inline bool sparse_worthwhile(int N, int n_l, int wG, const int* wLs) {
  if (wG < 1 || wG > SPARSE_MAX_WG) return false;
  long long cost = 2LL * wG;
  for (int l = 0; l < n_l; l++) {
    if (wLs[l] > SPARSE_MAX_WL) return false;
    cost += (long long)wLs[l] * wLs[l];
  }
  return cost * 6 <= (long long)N * (1 + 2 * n_l);
}

static inline bool sparse_plan(int N, int n_g, int n_l, int wG, int wL, int nnzE, int max_smem, SparsePlan* pl) {
  if (N < 2 || N > 0xffff) return false;
  pl->wG = wG; pl->wL = wL < 1 ? 1 : wL; pl->nnzE = nnzE;
  const long long n_el = (long long)N * (N + 1) / 2;
  pl->ld = N | 1;
  static const int epts[4] = {1, 2, 3, 4};
  pl->ept = 0;
  for (int q = 0; q < 4; q++) {
    long long th = ((n_el + epts[q] - 1) / epts[q] + 31) / 32 * 32;
    if (th <= (q == 0 ? 128 : (q <= 2 ? SPARSE_MAXT_KREG3 : SPARSE_MAXT_KREG))) { pl->ept = epts[q]; pl->threads = (int)th; pl->kreg = true; break; }
  }
  if (!pl->ept) {
    long long th = ((n_el + 7) / 8 + 31) / 32 * 32;
    if (th > SPARSE_MAXT_STREAM) return false;
    pl->ept = 8; pl->threads = (int)th; pl->kreg = false;
  }
  for (int yb = 2; yb >= 1; yb--) {
    pl->ybufs = yb;
    pl->smem = sparse_smem_bytes(N, pl->ld, yb, pl->wG, pl->wL, n_l, n_g);
    if (pl->smem <= (size_t)max_smem) return true;
  }
  return false;
}

static inline size_t sparse_ws_elems(const SparsePlan& pl) { return pl.kreg ? 0 : (size_t)6 * pl.ept * pl.threads; }

// pack buffer layout (bytes, 16-byte aligned pieces)
struct SparsePack {
  size_t o_gcol, o_g0, o_gk, o_lcol, o_lval, o_ij, o_eoff, o_eij, o_eval, bytes;
};
static inline SparsePack sparse_pack_layout(int N, int n_g, int n_l, int n_e, const SparsePlan& pl) {
  auto up = [](size_t x) { return (x + 15) / 16 * 16; };
  SparsePack k;
  size_t o = 0;
  k.o_gcol = o; o += up((size_t)pl.wG * N * sizeof(int));
  k.o_g0 = o; o += up((size_t)pl.wG * N * sizeof(cplx));
  k.o_gk = o; o += up((size_t)(n_g > 0 ? n_g : 1) * pl.wG * N * sizeof(cplx));
  k.o_lcol = o; o += up((size_t)(n_l > 0 ? n_l : 1) * pl.wL * N * sizeof(int));
  k.o_lval = o; o += up((size_t)(n_l > 0 ? n_l : 1) * pl.wL * N * sizeof(cplx));
  k.o_ij = o; o += up((size_t)N * (N + 1) / 2 * sizeof(unsigned));
  k.o_eoff = o; o += up((size_t)(n_e + 1) * sizeof(int));
  k.o_eij = o; o += up((size_t)(pl.nnzE > 0 ? pl.nnzE : 1) * sizeof(unsigned));
  k.o_eval = o; o += up((size_t)(pl.nnzE > 0 ? pl.nnzE : 1) * sizeof(cplx));
  k.bytes = o + 64;
  return k;
}

template <int EPT, bool KREG>
static int launch_sparse_t(const SolveParams& p, const SparseOps& sp, const SparsePlan& pl, cudaStream_t st) {
  if (sp.wL == 1) {
    if (cudaFuncSetAttribute(solve_sparse_kernel<EPT, KREG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return SEQ_ERR_CUDA;
    solve_sparse_kernel<EPT, KREG, true><<<p.B, pl.threads, pl.smem, st>>>(p, sp);
  } else {
    if (cudaFuncSetAttribute(solve_sparse_kernel<EPT, KREG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem) != cudaSuccess) return SEQ_ERR_CUDA;
    solve_sparse_kernel<EPT, KREG, false><<<p.B, pl.threads, pl.smem, st>>>(p, sp);
  }
  return SEQ_OK;
}

static inline int launch_sparse_kernel(const SolveParams& p, const SparseOps& sp, const SparsePlan& pl, cudaStream_t st) {
  if (!pl.kreg) return launch_sparse_t<8, false>(p, sp, pl, st);
  switch (pl.ept) {
    case 1: return launch_sparse_t<1, true>(p, sp, pl, st);
    case 2: return launch_sparse_t<2, true>(p, sp, pl, st);
    case 3: return launch_sparse_t<3, true>(p, sp, pl, st);
    case 4: return launch_sparse_t<4, true>(p, sp, pl, st);
  }
  return SEQ_ERR_UNSUPPORTED;
}

}  // namespace seq
