// Hermitian-tile stepper of the large path (kernel_large_herm.cuh) with the OUTERMOST tensor factor kept inside a thread.
//
// The reference's large systems are tensor products whose first mode is small (C5: transmon(3) x cavity(25) x
// cavity(25), system.py:399-445): with i = q L + x (L = N / Q, Q <= 4) every operator of that mode is a diagonal at an
// offset that is a multiple of L.  A thread owns the Q x Q elements { rho[q L + x, q' L + x'] } of ONE pair (x, x'):
// the terms at offsets dq L - the qubit drive's four gathers, the qubit decay's gather: five of the fifteen gathers of
// a C5 element - read the thread's own registers (what solve_diagblk_kernel does for the small systems), and the
// remaining gathers (offsets that are not multiples of L) go to the full stage matrix as before.  Nothing is assumed
// about the DATA: a term is local because of its OFFSET (Y[i + dq L, j] is element (q + dq, q') of the same thread
// whenever it lies inside the matrix), so any operator set whose offsets happen to contain multiples of a divisor L of N
// takes this path and all others take the 32 x 32 tiles of kernel_large_herm.cuh.
//
// Tiles are 16 x 16 pairs (x, x') of the x-space (tile pairs xi <= xj; 820 CTAs of 256 threads at C5); packed vectors
// are [tile pair][q Q + q'][16 x 16].  Positions x >= L of the last x tile have zero table entries (their elements come
// out as exact zeros) and are never written into the full matrix - there they would be rows of the next q block.
#pragma once
#include "kernel_large_herm.cuh"

namespace seq {

#define LQ_T 16
#define LQ_TE (LQ_T * LQ_T)
#define LQ_GLOBAL 99      // term is not local: gathered from the full stage matrix

struct LargeQLocal {
  int L, Q;
  int go_dq[DIAG_MAX_GD];      // per off-diagonal G term: offset / L, or LQ_GLOBAL
  int lg_dq[DIAG_MAX_LT];      // per gathered collapse pair: o_a / L = o_b / L, or LQ_GLOBAL
};

template <int Q>
static inline size_t lhq_stage_smem(int nD, int nT) {
  return ((size_t)Q * Q * LQ_T * (LQ_T + 1) + (size_t)(nD + nT + 1) * 2 * Q * LQ_T) * sizeof(cplx) + 40 * sizeof(double);
}

// v[q Q + q']: element (q L + x, q' L + x') of this thread's pair (x, x') = (xi 16 + px, xj 16 + py) -> the full matrix,
// and (xi != xj) the conjugated mirrors (q' L + x', q L + x)
template <int Q>
__device__ __forceinline__ void lhq_write(cplx* __restrict__ Yo, int NP, int L, int xi, int xj, const cplx (&v)[Q * Q], cplx* __restrict__ sm) {
  const int px = threadIdx.x >> 4, py = threadIdx.x & 15;
  const int x = xi * LQ_T + px, x2 = xj * LQ_T + py;
  if (x < L && x2 < L) {
    cplx* po = Yo + (long long)x * NP + x2;
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) po[(long long)q * L * NP + q2 * L] = v[q * Q + q2];
  }
  if (xi != xj) {      // (block-uniform)
#pragma unroll
    for (int e = 0; e < Q * Q; e++) sm[e * (LQ_T * (LQ_T + 1)) + px * (LQ_T + 1) + py] = v[e];
    __syncthreads();
    // thread (px, py) writes the mirror element of row x' = xj 16 + px, column x = xi 16 + py
    const int mr = xj * LQ_T + px, mc = xi * LQ_T + py;
    if (mr < L && mc < L) {
      cplx* pm = Yo + (long long)mr * NP + mc;
#pragma unroll
      for (int q = 0; q < Q; q++)
#pragma unroll
        for (int q2 = 0; q2 < Q; q2++)      // mirror of (q, x = mc; q2, x' = mr) is row (q2, mr), column (q, mc)
          pm[(long long)q2 * L * NP + q * L] = cconj(sm[(q * Q + q2) * (LQ_T * (LQ_T + 1)) + py * (LQ_T + 1) + px]);
    }
  }
}

template <int Q>
__global__ void __launch_bounds__(256) lhq_pack_kernel(const cplx* __restrict__ src, const __grid_constant__ LargeHermBufs bf, int L,
                                                       unsigned long long* __restrict__ chk) {
  extern __shared__ __align__(16) unsigned char lh_smem[];
  cplx* sm = (cplx*)lh_smem;
  __shared__ double red[16];
  const unsigned code = bf.tiles[blockIdx.x];
  const int xi = code & 0xffffu, xj = code >> 16, N = bf.N;
  const int px = threadIdx.x >> 4, py = threadIdx.x & 15, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int x = xi * LQ_T + px, x2 = xj * LQ_T + py;
  const bool valid = x < L && x2 < L;
  cplx* yp = bf.ybuf[0] + (long long)blockIdx.x * (Q * Q * LQ_TE) + threadIdx.x;
  cplx v[Q * Q];
  double mx = 0.0, as = 0.0;
#pragma unroll
  for (int q = 0; q < Q; q++)
#pragma unroll
    for (int q2 = 0; q2 < Q; q2++) {
      cplx t = cmake(0, 0);
      if (valid) {
        const long long i = q * L + x, j = q2 * L + x2;
        t = src[i * N + j];
        const cplx m = src[j * N + i];
        mx = fmax(mx, cabs2(t));
        const double dx = m.x - t.x, dy = m.y + t.y;
        as = fmax(as, dx * dx + dy * dy);
      }
      v[q * Q + q2] = t;
      yp[(q * Q + q2) * LQ_TE] = t;
    }
  lhq_write<Q>(bf.Yfull[0], bf.NPh, L, xi, xj, v, sm);
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); as = fmax(as, __shfl_xor_sync(0xffffffffu, as, o)); }
  if (lane == 0) { red[w] = mx; red[8 + w] = as; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) { mx = fmax(mx, red[q]); as = fmax(as, red[8 + q]); }
    atomicMax(&chk[0], (unsigned long long)__double_as_longlong(mx));
    atomicMax(&chk[1], (unsigned long long)__double_as_longlong(as));
  }
}

// stage input of the attempt's first stage (see lh_first_kernel); blocks behind the tiles: expectation values
template <int Q>
__global__ void __launch_bounds__(256) lhq_first_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int L, int n_tiles,
                                                        const __grid_constant__ LargeHermExpect ex) {
  extern __shared__ __align__(16) unsigned char lh_smem[];
  cplx* sm = (cplx*)lh_smem;
  lh_grid_dependency_wait();
  if ((int)blockIdx.x >= n_tiles) {
    lh_expect_body(dev, (int)blockIdx.x - n_tiles, ex, bf.Yfull[0], bf.NPh, (double*)sm);
    return;
  }
  if (!dev->active) return;
  if (dev->st0 == 0) return;
  const unsigned code = bf.tiles[blockIdx.x];
  const int xi = code & 0xffffu, xj = code >> 16;
  const long long tbase = (long long)blockIdx.x * (Q * Q * LQ_TE) + threadIdx.x;
  const cplx* y = bf.ybuf[dev->ycur] + tbase;
  const cplx* k0 = bf.kbuf[dev->kmap[0]] + tbase;
  const double h = dev->htry, a10 = RK_A[dev->ctl.pair][1][0];
  cplx v[Q * Q];
#pragma unroll
  for (int e = 0; e < Q * Q; e++) {
    const cplx yv = LQ_LD(y + e * LQ_TE), kv = LQ_LD(k0 + e * LQ_TE);
    v[e] = cmake(fma(h, fma(a10, kv.x, 0.0), yv.x), fma(h, fma(a10, kv.y, 0.0), yv.y));
  }
  lhq_write<Q>(bf.Yfull[1], bf.NPh, L, xi, xj, v, sm);
}

// terms at offset DQ L: sources are this thread's own elements
template <int Q, int DQ>
__device__ __forceinline__ void lhq_local_go(cplx (&kv)[Q * Q], const cplx (&yv)[Q * Q], const cplx* __restrict__ trow, const cplx* __restrict__ tcol) {
#pragma unroll
  for (int q = 0; q < Q; q++) {
    if (q + DQ < 0 || q + DQ >= Q) continue;
    const cplx vi = trow[q * LQ_T];
    const cplx mvi = cmake(vi.y, -vi.x);                                 // -i v_d[i]
#pragma unroll
    for (int q2 = 0; q2 < Q; q2++) cfma(kv[q * Q + q2], mvi, yv[(q + DQ) * Q + q2]);
  }
#pragma unroll
  for (int q2 = 0; q2 < Q; q2++) {
    if (q2 + DQ < 0 || q2 + DQ >= Q) continue;
    const cplx vj = tcol[q2 * LQ_T];
    const cplx c = cmake(vj.y, vj.x);                                    // +i conj(v_d[j])
#pragma unroll
    for (int q = 0; q < Q; q++) cfma(kv[q * Q + q2], c, yv[q * Q + q2 + DQ]);
  }
}
template <int Q, int DQ>
__device__ __forceinline__ void lhq_local_lg(cplx (&kv)[Q * Q], const cplx (&yv)[Q * Q], const cplx* __restrict__ larow, const cplx* __restrict__ lbcol, double rr) {
#pragma unroll
  for (int q = 0; q < Q; q++) {
    if (q + DQ < 0 || q + DQ >= Q) continue;
    const cplx la = larow[q * LQ_T];
#pragma unroll
    for (int q2 = 0; q2 < Q; q2++) {
      if (q2 + DQ < 0 || q2 + DQ >= Q) continue;
      const cplx lb = lbcol[q2 * LQ_T];
      const cplx wv = cmake(rr * (la.x * lb.x + la.y * lb.y), rr * (la.y * lb.x - la.x * lb.y));
      cfma(kv[q * Q + q2], wv, yv[(q + DQ) * Q + q2 + DQ]);
    }
  }
}

template <int Q>
__global__ void __launch_bounds__(256, (Q > 3 ? 1 : 2)) lhq_stage_kernel(LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int st,
                                                           const __grid_constant__ LargeDiagArgs a, const __grid_constant__ LargeQLocal lq, double atol, double rtol) {
  extern __shared__ __align__(16) unsigned char lh_smem[];
  lh_grid_dependency_wait();
  if (!dev->active || st < dev->st0 || st >= dev->S) return;
  constexpr int E = Q * Q, TS = 2 * Q * LQ_T;      // elements per thread; entries of one tile table (rows of every q, then columns)
  cplx* sm = (cplx*)lh_smem;                                // transpose buffers E x 16 x 17
  cplx* gts = sm + E * LQ_T * (LQ_T + 1);                   // [nD][TS]: G(t) diagonals of this stage time
  cplx* lams = gts + bf.nD * TS;                            // [nT][TS]: collapse diagonals
  cplx* gds = lams + bf.nT * TS;                            // [TS]: static main diagonal of G0
  double* red = (double*)(gds + TS);
  const int pair = dev->ctl.pair, S = dev->S, NP = bf.NPh, L = lq.L;
  const bool last = st == S - 1;
  const double h = dev->htry;
  const unsigned code = bf.tiles[blockIdx.x];
  const int xi = code & 0xffffu, xj = code >> 16;
  const int px = threadIdx.x >> 4, py = threadIdx.x & 15;
  const long long tbase = (long long)blockIdx.x * (E * LQ_TE) + threadIdx.x;
  const double* cval = dev->cval + st * LDEV_MAXG;

  // ---- tile tables: entries of positions x >= L are zero (those elements then come out as exact zeros)
  for (int idx = threadIdx.x; idx < (bf.nD + bf.nT + 1) * TS; idx += blockDim.x) {
    const int d = idx / TS, rem = idx - d * TS;
    const int side = rem / (Q * LQ_T), q = (rem / LQ_T) % Q, xx = (side ? xj : xi) * LQ_T + (rem % LQ_T);
    cplx g = cmake(0, 0);
    if (xx < L) {
      const int pos = q * L + xx;
      if (d < bf.nD) {      // g0 + sum_m c_m(t) gk_m
        g = bf.g0[(long long)d * bf.N + pos];
        for (int m = 0; m < bf.n_g; m++) {
          const cplx gm = bf.gk[((long long)m * bf.nD + d) * bf.N + pos];
          const double c = cval[m];
          g.x = fma(c, gm.x, g.x);
          g.y = fma(c, gm.y, g.y);
        }
      } else if (d < bf.nD + bf.nT) {
        g = bf.lam[(long long)(d - bf.nD) * bf.N + pos];
      } else {
        g = bf.gdiag[pos];
      }
    }
    gts[idx] = g;
  }
  __syncthreads();

  // ---- right-hand side of the Q x Q elements of pair (x, x')
  const cplx* __restrict__ Y = bf.Yfull[st & 1];
  const cplx* p0 = Y + (long long)(xi * LQ_T + px) * NP + xj * LQ_T + py;      // element (q = 0, q' = 0)
  const long long LR = (long long)L * NP;
  const int ro = px, co = Q * LQ_T + py;      // row / column entry of q = 0 inside a tile table
  cplx yv[E], kv[E];
#pragma unroll
  for (int q = 0; q < Q; q++)
#pragma unroll
    for (int q2 = 0; q2 < Q; q2++) yv[q * Q + q2] = p0[q * LR + q2 * L];
  {
    cplx gi[Q], gj[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) { gi[q] = gds[ro + q * LQ_T]; gj[q] = gds[co + q * LQ_T]; }
#pragma unroll
    for (int q = 0; q < Q; q++)
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) kv[q * Q + q2] = cmul(cmake(gi[q].y + gj[q2].y, -(gi[q].x - gj[q2].x)), yv[q * Q + q2]);      // -i (G0_ii - conj(G0_jj)) y
  }
  for (int t = 0; t < a.nLs; t++) lhq_local_lg<Q, 0>(kv, yv, lams + a.ls_a[t] * TS + ro, lams + a.ls_b[t] * TS + co, 1.0);
  if (a.g0_tab >= 0) {
    const cplx* gv = gts + a.g0_tab * TS;
#pragma unroll
    for (int q = 0; q < Q; q++) {
      const cplx vi = gv[ro + q * LQ_T];
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) {
        const cplx vj = gv[co + q2 * LQ_T];
        cfma(kv[q * Q + q2], cmake(vi.y + vj.y, -(vi.x - vj.x)), yv[q * Q + q2]);
      }
    }
  }
  for (int t = 0; t < a.nLo; t++) lhq_local_lg<Q, 0>(kv, yv, lams + a.lo_a[t] * TS + ro, lams + a.lo_b[t] * TS + co, cval[a.lo_r[t]]);
  for (int t = 0; t < a.nGo; t++) {
    const cplx* gv = gts + a.go_tab[t] * TS;
    const int dq = lq.go_dq[t];
    if (dq == LQ_GLOBAL) {
      const int o = a.go_off[t];
      const cplx* pr = p0 + (long long)o * NP;                   // Y[i + o, j]
      const cplx* pc = p0 + o;                                   // Y[i, j + o]
      cplx cj[Q];
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) { const cplx vj = gv[co + q2 * LQ_T]; cj[q2] = cmake(vj.y, vj.x); }
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const cplx vi = gv[ro + q * LQ_T];
        const cplx mvi = cmake(vi.y, -vi.x);
#pragma unroll
        for (int q2 = 0; q2 < Q; q2++) {
          const cplx ya = pr[q * LR + q2 * L], yb = pc[q * LR + q2 * L];
          cfma(kv[q * Q + q2], mvi, ya);
          cfma(kv[q * Q + q2], cj[q2], yb);
        }
      }
    } else {
      switch (dq) {
        case 1: lhq_local_go<Q, 1>(kv, yv, gv + ro, gv + co); break;
        case -1: lhq_local_go<Q, -1>(kv, yv, gv + ro, gv + co); break;
        case 2: if (Q > 2) lhq_local_go<Q, (Q > 2 ? 2 : 1)>(kv, yv, gv + ro, gv + co); break;
        case -2: if (Q > 2) lhq_local_go<Q, (Q > 2 ? -2 : -1)>(kv, yv, gv + ro, gv + co); break;
        case 3: if (Q > 3) lhq_local_go<Q, (Q > 3 ? 3 : 1)>(kv, yv, gv + ro, gv + co); break;
        case -3: if (Q > 3) lhq_local_go<Q, (Q > 3 ? -3 : -1)>(kv, yv, gv + ro, gv + co); break;
        default: break;
      }
    }
  }
  for (int t = 0; t < a.nLg; t++) {
    const int ri = a.lg_r[t];
    const double rr = ri < 0 ? 1.0 : cval[ri];
    const cplx* larow = lams + a.lg_a[t] * TS + ro;
    const cplx* lbcol = lams + a.lg_b[t] * TS + co;
    const int dq = lq.lg_dq[t];
    if (dq == LQ_GLOBAL) {
      const cplx* pg = p0 + (long long)a.lg_oa[t] * NP + a.lg_ob[t];
      cplx lb[Q];
#pragma unroll
      for (int q2 = 0; q2 < Q; q2++) lb[q2] = lbcol[q2 * LQ_T];
#pragma unroll
      for (int q = 0; q < Q; q++) {
        const cplx la = larow[q * LQ_T];
#pragma unroll
        for (int q2 = 0; q2 < Q; q2++) {
          const cplx wv = cmake(rr * (la.x * lb[q2].x + la.y * lb[q2].y), rr * (la.y * lb[q2].x - la.x * lb[q2].y));
          cfma(kv[q * Q + q2], wv, pg[q * LR + q2 * L]);
        }
      }
    } else {
      switch (dq) {
        case 1: lhq_local_lg<Q, 1>(kv, yv, larow, lbcol, rr); break;
        case -1: lhq_local_lg<Q, -1>(kv, yv, larow, lbcol, rr); break;
        case 2: if (Q > 2) lhq_local_lg<Q, (Q > 2 ? 2 : 1)>(kv, yv, larow, lbcol, rr); break;
        case -2: if (Q > 2) lhq_local_lg<Q, (Q > 2 ? -2 : -1)>(kv, yv, larow, lbcol, rr); break;
        case 3: if (Q > 3) lhq_local_lg<Q, (Q > 3 ? 3 : 1)>(kv, yv, larow, lbcol, rr); break;
        case -3: if (Q > 3) lhq_local_lg<Q, (Q > 3 ? -3 : -1)>(kv, yv, larow, lbcol, rr); break;
        default: break;
      }
    }
  }
  {
    cplx* kd = bf.kbuf[dev->kmap[st]] + tbase;
#pragma unroll
    for (int e = 0; e < E; e++) LQ_ST(kd + e * LQ_TE, kv[e]);
  }
  const cplx* y = bf.ybuf[dev->ycur] + tbase;
  if (!last) {
    // Y_{st+1} = y + h sum_{jj <= st} a_{st+1,jj} k_jj   (zero coefficients skipped: fma(0, k, acc) = acc)
    cplx acc[E];
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = cmake(0, 0);
    for (int jj = 0; jj < st; jj++) {
      const double aj = RK_A[pair][st + 1][jj];
      if (aj == 0.0) continue;
      const cplx* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const cplx kq = LQ_LD(kj + e * LQ_TE);
        acc[e].x = fma(aj, kq.x, acc[e].x);
        acc[e].y = fma(aj, kq.y, acc[e].y);
      }
    }
    const double as = RK_A[pair][st + 1][st];
#pragma unroll
    for (int e = 0; e < E; e++) {
      const cplx yq = LQ_LD(y + e * LQ_TE);
      acc[e].x = fma(as, kv[e].x, acc[e].x);
      acc[e].y = fma(as, kv[e].y, acc[e].y);
      acc[e] = cmake(fma(h, acc[e].x, yq.x), fma(h, acc[e].y, yq.y));
    }
    if (st == S - 2) {      // the propagated solution: the new state if the step is accepted
      cplx* yn = bf.ybuf[dev->ycur ^ 1] + tbase;
#pragma unroll
      for (int e = 0; e < E; e++) LQ_ST(yn + e * LQ_TE, acc[e]);
    }
    lhq_write<Q>(bf.Yfull[(st + 1) & 1], NP, L, xi, xj, acc, sm);
  } else {
    // error estimate (see lh_stage_kernel); the stage input of the last stage (yv) is y_new
    cplx er[E];
#pragma unroll
    for (int e = 0; e < E; e++) er[e] = cmake(0, 0);
    for (int jj = 0; jj < S - 1; jj++) {
      const double ej = RK_E[pair][jj];
      if (ej == 0.0) continue;
      const cplx* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const cplx kq = LQ_LD(kj + e * LQ_TE);
        er[e].x = fma(ej, kq.x, er[e].x);
        er[e].y = fma(ej, kq.y, er[e].y);
      }
    }
    const double el = RK_E[pair][S - 1];
    const double m2floor = atol >= 1e-100 ? 1e-280 : 0.0;
    double es = 0.0;
    if (xi * LQ_T + px < L && xj * LQ_T + py < L) {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const double ex = fma(el, kv[e].x, er[e].x) * h, ey = fma(el, kv[e].y, er[e].y) * h;
        const cplx yq = LQ_LD(y + e * LQ_TE);
        const double s = atol + rtol * sqrt(fmax(fmax(yq.x * yq.x + yq.y * yq.y, yv[e].x * yv[e].x + yv[e].y * yv[e].y), m2floor));
        es = fma(ex * ex + ey * ey, 1.0 / (s * s), es);
      }
    }
    es = block_sum(es, red);
    if (threadIdx.x == 0) atomicAdd(&dev->err2, xi == xj ? es : 2.0 * es);
  }
}

// Largest offset L among the gathered terms with N = Q L, Q in {2, 3, 4}; fills the per-term classification.
static inline bool lhq_plan(int N, const LargeDiagArgs& a, LargeQLocal* lq) {
  memset(lq, 0, sizeof(*lq));
  int best = 0;
  auto consider = [&](int o) {
    o = abs(o);
    if (o > best && o >= LQ_T && N % o == 0 && N / o >= 2 && N / o <= 4) best = o;
  };
  for (int t = 0; t < a.nGo; t++) consider(a.go_off[t]);
  for (int t = 0; t < a.nLg; t++) { consider(a.lg_oa[t]); consider(a.lg_ob[t]); }
  if (!best) return false;
  lq->L = best; lq->Q = N / best;
  for (int t = 0; t < a.nGo; t++) lq->go_dq[t] = (a.go_off[t] % best == 0) ? a.go_off[t] / best : LQ_GLOBAL;
  for (int t = 0; t < a.nLg; t++)
    lq->lg_dq[t] = (a.lg_oa[t] == a.lg_ob[t] && a.lg_oa[t] % best == 0 && a.lg_oa[t] != 0) ? a.lg_oa[t] / best : LQ_GLOBAL;
  return true;
}

}  // namespace seq
