// Structured right-hand side of the large-system path (SEQ_METHOD_LARGE, C5: transmon (x) cavity (x) cavity,
// N = 1875): when every operator is a sum of a few diagonals (kernel_diag.cuh has the argument: the
// reference's operators are Kronecker-embedded ladder / number operators) ONE grid-wide kernel evaluates
//     K_ij = W_ij Y_ij - i sum_d v_d[i](t) Y[i+o_d, j] + i sum_d conj(v_d[j](t)) Y[i, j+o_d]
//            + sum_t r_t lam_a[i] conj(lam_b[j]) Y[i+o_a, j+o_b]
// for this rank's rows straight from the planar stage state - a few dozen multiply-adds and (cached) gathers
// per element instead of (2 + 2 n_l) N^3-flop ZGEMMs; the dense operators are never materialised (0.9 GB at
// C5).  HBM-bound: per RHS the state is read once from L2/HBM (the shifted gathers hit L1/L2) and K written
// once; algorithmic bytes 32 N^2 per RHS (16 read + 16 written).
#pragma once
#include "common.cuh"
#include "kernel_diag.cuh"

namespace seq {

struct LargeDiagArgs {
  int N, NP, m, Mrows;
  long long row0, y_plane, k_plane;
  const double* Y;         // full stage state, planar [2][rows_full][NP]
  double* K;               // this rank's rows of the derivative, planar [2][m][NP]
  const cplx* gt;          // [nD][N]: G diagonals at this stage time (static main diagonal excluded)
  const cplx* lam;         // [nT][N]: collapse diagonals
  const cplx* gdiag;       // [N]: static main diagonal of G0
  const double* cval;      // spline values of this stage time
  int g0_tab;              // table of the time-dependent main diagonal, or -1
  int nGo, go_tab[DIAG_MAX_GD], go_off[DIAG_MAX_GD];
  int nLs, ls_a[DIAG_MAX_LT], ls_b[DIAG_MAX_LT];
  int nLo, lo_a[DIAG_MAX_LT], lo_b[DIAG_MAX_LT], lo_r[DIAG_MAX_LT];
  int nLg, lg_a[DIAG_MAX_LT], lg_b[DIAG_MAX_LT], lg_r[DIAG_MAX_LT], lg_oa[DIAG_MAX_LT], lg_ob[DIAG_MAX_LT];
};

// gt[q] = g0[q] + sum_m cval[m] gk[m][q]   (q over nD * N diagonal entries)
__global__ void large_diag_gt_kernel(const cplx* __restrict__ g0, const cplx* __restrict__ gk, const double* __restrict__ cval,
                                     int n_g, int n, cplx* __restrict__ gt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  cplx g = g0[q];
  for (int m = 0; m < n_g; m++) {
    const cplx a = gk[(long long)m * n + q];
    const double v = cval[m];
    g.x = fma(v, a.x, g.x);
    g.y = fma(v, a.y, g.y);
  }
  gt[q] = g;
}

// one element K_ij of the structured right-hand side from the planar stage state (Yr / Yi: real / imaginary plane, leading
// dimension NP); i, j < N
// G(t) diagonal `tab` at this element's row / column: from the table in global memory ...
struct LargeGtGlobal {
  const cplx* gt; int N; long long i; int j;
  __device__ __forceinline__ cplx row(int tab) const { return gt[(long long)tab * N + i]; }
  __device__ __forceinline__ cplx col(int tab) const { return gt[(long long)tab * N + j]; }
};
// ... or from the 2 x 32 entries a tile kernel built in shared memory ([tab][row | column][32])
struct LargeGtTile {
  const cplx* gts; int r, c;
  __device__ __forceinline__ cplx row(int tab) const { return gts[tab * 64 + r]; }
  __device__ __forceinline__ cplx col(int tab) const { return gts[tab * 64 + 32 + c]; }
};

template <class GT>
__device__ __forceinline__ cplx large_diag_rhs_elem(const LargeDiagArgs& a, const double* __restrict__ Yr, const double* __restrict__ Yi, int NP,
                                                    const double* __restrict__ cval, long long i, int j, const GT& G) {
  const int N = a.N;
  const cplx y = cmake(Yr[i * NP + j], Yi[i * NP + j]);
  // static own-element weight
  const cplx gi = a.gdiag[i], gj = a.gdiag[j];
  cplx w = cmake(gi.y + gj.y, -(gi.x - gj.x));
  for (int t = 0; t < a.nLs; t++) {
    const cplx la = a.lam[(long long)a.ls_a[t] * N + i], lb = a.lam[(long long)a.ls_b[t] * N + j];
    w.x += la.x * lb.x + la.y * lb.y;
    w.y += la.y * lb.x - la.x * lb.y;
  }
  if (a.g0_tab >= 0) {
    const cplx vi = G.row(a.g0_tab), vj = G.col(a.g0_tab);
    w.x += vi.y + vj.y;
    w.y -= vi.x - vj.x;
  }
  for (int t = 0; t < a.nLo; t++) {
    const double rr = cval[a.lo_r[t]];
    const cplx la = a.lam[(long long)a.lo_a[t] * N + i], lb = a.lam[(long long)a.lo_b[t] * N + j];
    w.x += rr * (la.x * lb.x + la.y * lb.y);
    w.y += rr * (la.y * lb.x - la.x * lb.y);
  }
  cplx acc = cmul(w, y);
  for (int t = 0; t < a.nGo; t++) {
    const int o = a.go_off[t];
    const long long ii = i + o;
    if (ii >= 0 && ii < N) {
      const cplx vi = G.row(a.go_tab[t]);
      const cplx ya = cmake(Yr[ii * NP + j], Yi[ii * NP + j]);
      cfma(acc, cmake(vi.y, -vi.x), ya);                       // -i v_d[i] Y[i+o, j]
    }
    const int jj = j + o;
    if (jj >= 0 && jj < N) {
      const cplx vj = G.col(a.go_tab[t]);
      const cplx yv = cmake(Yr[i * NP + jj], Yi[i * NP + jj]);
      cfma(acc, cmake(vj.y, vj.x), yv);                        // +i conj(v_d[j]) Y[i, j+o]
    }
  }
  for (int t = 0; t < a.nLg; t++) {
    const long long ii = i + a.lg_oa[t];
    const int jj = j + a.lg_ob[t];
    if (ii >= 0 && ii < N && jj >= 0 && jj < N) {
      const int ri = a.lg_r[t];
      const double rr = ri < 0 ? 1.0 : cval[ri];
      const cplx la = a.lam[(long long)a.lg_a[t] * N + i], lb = a.lam[(long long)a.lg_b[t] * N + j];
      const cplx wv = cmake(rr * (la.x * lb.x + la.y * lb.y), rr * (la.y * lb.x - la.x * lb.y));
      cfma(acc, wv, cmake(Yr[ii * NP + jj], Yi[ii * NP + jj]));
    }
  }
  return acc;
}

// one thread per element of this rank's row block (padding included: it is written as zero)
// (K and cval as separate arguments: the device-controlled stepper picks them per attempt; a stays in the constant bank)
__device__ __forceinline__ void large_diag_rhs_body(const LargeDiagArgs& a, double* __restrict__ Kout, const double* __restrict__ cval) {
  const int N = a.N, NP = a.NP;
  const long long total = (long long)a.m * NP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / NP), j = (int)(idx - (long long)r * NP);
    const long long i = a.row0 + r;
    cplx acc = cmake(0, 0);
    if (r < a.Mrows && i < N && j < N) acc = large_diag_rhs_elem(a, a.Y, a.Y + a.y_plane, NP, cval, i, j, LargeGtGlobal{a.gt, N, i, j});
    Kout[idx] = acc.x;
    Kout[a.k_plane + idx] = acc.y;
  }
}
__global__ void __launch_bounds__(256) large_diag_rhs_kernel(const __grid_constant__ LargeDiagArgs a) { large_diag_rhs_body(a, a.K, a.cval); }

// term lists in element units from the host copy of diag_flags_kernel's output (same walk as diag_plan)
static inline bool large_diag_terms(int N, int n_c, int n_ctd, int n_ch, const int* flags, LargeDiagArgs* a, DiagFill* f) {
  const int n_l = n_c + n_ctd, W = 2 * N - 1;
  memset(f, 0, sizeof(*f));
  a->g0_tab = -1; a->nGo = 0;
  int nD = 0;
  for (int q = 0; q < W; q++) {
    if (!flags[q]) continue;
    const int off = q - (N - 1);
    if (off == 0 && !flags[(long long)(1 + n_l) * W + q]) continue;
    if (nD >= DIAG_MAX_GD) return false;
    f->goff[nD] = off;
    if (off == 0) a->g0_tab = nD;
    else { a->go_tab[a->nGo] = nD; a->go_off[a->nGo] = off; a->nGo++; }
    nD++;
  }
  if (nD == 0) { f->goff[0] = 0; nD = 1; }
  f->nD = nD;
  int nT = 0;
  a->nLs = a->nLo = a->nLg = 0;
  for (int l = 0; l < n_l; l++) {
    const int first = nT;
    for (int q = 0; q < W; q++) {
      if (!flags[(long long)(1 + l) * W + q]) continue;
      if (nT >= DIAG_MAX_TAB) return false;
      f->t_op[nT] = (short)l; f->t_off[nT] = q - (N - 1);
      nT++;
    }
    if (nT - first > 4) return false;
    const int rate = (l < n_c) ? -1 : (n_ch + (l - n_c));
    for (int x = first; x < nT; x++)
      for (int y = first; y < nT; y++) {
        if (f->t_off[x] == 0 && f->t_off[y] == 0) {
          if (a->nLs >= DIAG_MAX_LT || a->nLo >= DIAG_MAX_LT) return false;
          if (rate < 0) { a->ls_a[a->nLs] = x; a->ls_b[a->nLs] = y; a->nLs++; }
          else { a->lo_a[a->nLo] = x; a->lo_b[a->nLo] = y; a->lo_r[a->nLo] = rate; a->nLo++; }
        } else {
          if (a->nLg >= DIAG_MAX_LT) return false;
          a->lg_a[a->nLg] = x; a->lg_b[a->nLg] = y; a->lg_r[a->nLg] = rate;
          a->lg_oa[a->nLg] = f->t_off[x]; a->lg_ob[a->nLg] = f->t_off[y];
          a->nLg++;
        }
      }
  }
  f->nT = nT;
  return true;
}

}  // namespace seq
