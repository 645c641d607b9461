// Large path (C5), Hermitian tiles: the device-controlled stepper of kernel_large_dev.cuh on HALF of rho, with the stage
// combination and the error estimate fused into the right-hand-side kernel.
//
// rho is Hermitian and so is every stage derivative (the Lindblad right-hand side maps Hermitian matrices to Hermitian
// matrices), so only the 32 x 32 tiles (ti <= tj) of the upper block triangle are integrated.  State and stage
// derivatives live PACKED ([tile][32 x 32] complex, half the bytes of a full vector); the stage input alone is kept
// as a full matrix (two of them, ping-pong), because the structured right-hand side gathers Y[i + o, j + o'] from
// anywhere.  One kernel per stage and tile does
//     k_st   = f(t_st, Y_st)                        gathers from the full stage matrix (L1 / L2)
//     Y_st+1 = y + h sum_{j <= st} a_{st+1,j} k_j    packed streams (HBM), k_st from registers
// and writes Y_st+1 into the other full matrix: its own tile and - through a shared-memory transpose, conjugated - the
// mirrored tile.  The last stage's epilogue is the error estimate instead (weight 2 off the block diagonal).  Per 4(3)
// step that is 4 fused kernels + 1 stage-input kernel over ~36 half vectors instead of 4 + 4 + 1 kernels over ~68;
// the arithmetic of an element (order of the multiply-adds) is the one of ldev_combine_kernel / ldev_err_kernel.
// Stage st reads full matrix st & 1 and writes (st + 1) & 1; both pairs have an even last stage, so the committed
// state always ends in matrix 0, where the outputs read it.
//
// Layout choices that shape the instruction stream (ncu on the first version: 550 warp instructions per element, issue
// slots half used, L2 and HBM a fifth / a third - the kernel waited on its own address and bounds arithmetic):
//  * full matrices INTERLEAVED complex [NPh][NPh] (one 16-byte load per gather instead of two 8-byte ones), with
//    `guard` zero rows above and below: a gather whose source lies outside the matrix has a ZERO coefficient by
//    construction of the diagonal tables (v_d[i] = G[i][i+o] exists only inside the matrix), so no gather is bounds
//    checked - it reads zeros or finite neighbours and multiplies them by zero.  Padding rows / columns (N..NPh) are
//    written as zeros by every stage for the same reason;
//  * every table value an element needs (G(t) diagonals - built here from the spline values -, collapse diagonals, the
//    static main diagonal) is staged per tile: 32 row entries + 32 column entries per table in shared memory;
//  * a thread owns 4 rows of one column: terms are the OUTER loop, so a term's parameters, its column-side coefficient and
//    its source offset are computed once per 4 elements.
#pragma once
#include "kernel_large_dev.cuh"

namespace seq {

#define LH_T 32
#define LH_TE (LH_T * LH_T)
// packed streams (state, stage derivatives) pass through once per kernel: streaming loads / stores keep L1 / L2 for the gathers
#ifdef SEQ_LQ_PLAIN_STREAMS
#define LQ_LD(p) (*(p))
#define LQ_ST(p, v) (*(p) = (v))
#else
#define LQ_LD(p) __ldcs(p)
#define LQ_ST(p, v) __stcs((p), (v))
#endif

struct LargeHermBufs {
  cplx* ybuf[2];            // packed state / new state
  cplx* kbuf[7];            // packed stage derivatives
  cplx* Yfull[2];           // full stage matrices [NPh][NPh] (pointers to element (0, 0); `guard` zero rows on either side)
  const unsigned* tiles;    // ti | tj << 16 of every upper tile
  int NPh, N;
  const cplx *g0, *gk;      // G diagonals: static part [nD][N] and time-dependent terms [n_g][nD][N] (diag_fill_kernel's tables)
  const cplx *lam, *gdiag;  // collapse diagonals [nT][N], static main diagonal of G0 [N]
  int n_g, nD, nT;
  int dbg;                  // timing experiments (SEQ_LH_DBG; results are wrong): 1 no mirror write, 2 no tile write, 4 no k store
};

static inline size_t lh_stage_smem(int nD, int nT) { return ((size_t)LH_T * (LH_T + 1) + (size_t)(nD + nT + 1) * 2 * LH_T) * sizeof(cplx) + 40 * sizeof(double); }

// v[q]: element (row w + 8 q, column lane) of tile (ti, tj) -> the tile and (ti != tj) its conjugate-transposed mirror
__device__ __forceinline__ void lh_write_tile(cplx* __restrict__ Yo, int NP, int ti, int tj, const cplx (&v)[4], cplx* __restrict__ sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  cplx* po = Yo + (long long)(ti * LH_T + w) * NP + tj * LH_T + lane;
#pragma unroll
  for (int q = 0; q < 4; q++) po[(long long)(8 * q) * NP] = v[q];
  if (ti != tj) {      // (block-uniform)
#pragma unroll
    for (int q = 0; q < 4; q++) sm[(w + 8 * q) * (LH_T + 1) + lane] = v[q];
    __syncthreads();
    cplx* pm = Yo + (long long)(tj * LH_T + w) * NP + ti * LH_T + lane;
#pragma unroll
    for (int q = 0; q < 4; q++) pm[(long long)(8 * q) * NP] = cconj(sm[lane * (LH_T + 1) + w + 8 * q]);
  }
}

// rho0 (interleaved [N][N]) -> packed state + full matrix 0; chk[0] = max |rho_ij|^2, chk[1] = max |rho_ji - conj(rho_ij)|^2
// as bit patterns of non-negative doubles (the host falls back to the full-matrix stepper for a non-Hermitian rho0)
__global__ void __launch_bounds__(256) lh_pack_kernel(const cplx* __restrict__ src, const __grid_constant__ LargeHermBufs bf, unsigned long long* __restrict__ chk) {
  __shared__ cplx sm[LH_T * (LH_T + 1)];
  __shared__ double red[16];
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16, N = bf.N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  cplx* yp = bf.ybuf[0] + (long long)blockIdx.x * LH_TE;
  cplx v[4];
  double mx = 0.0, as = 0.0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int r = w + 8 * q, i = ti * LH_T + r, j = tj * LH_T + lane;
    v[q] = cmake(0, 0);
    if (i < N && j < N) {
      v[q] = src[(long long)i * N + j];
      const cplx m = src[(long long)j * N + i];
      mx = fmax(mx, cabs2(v[q]));
      const double dx = m.x - v[q].x, dy = m.y + v[q].y;
      as = fmax(as, dx * dx + dy * dy);
    }
    yp[r * LH_T + lane] = v[q];
  }
  lh_write_tile(bf.Yfull[0], bf.NPh, ti, tj, v, sm);
  // block maxima (values are >= 0: the integer order of the bit patterns is the order of the numbers)
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); as = fmax(as, __shfl_xor_sync(0xffffffffu, as, o)); }
  if (lane == 0) { red[w] = mx; red[8 + w] = as; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) { mx = fmax(mx, red[q]); as = fmax(as, red[8 + q]); }
    atomicMax(&chk[0], (unsigned long long)__double_as_longlong(mx));
    atomicMax(&chk[1], (unsigned long long)__double_as_longlong(as));
  }
}

struct LargeHermExpect {
  const int* e_off; const unsigned* e_ij; const cplx* e_val;
  void* expect; long long base; int n_out, complex_out;
};
// Tr(E_e rho) of the committed state (full matrix 0) from the COO lists of the expectation operators
__device__ __forceinline__ void lh_expect_body(const LargeDev* __restrict__ dev, const int e, const LargeHermExpect& ex, const cplx* __restrict__ Y, int NP,
                                               double* red) {
  if (!dev->emit) return;
  cplx acc = cmake(0, 0);
  for (int q = ex.e_off[e] + threadIdx.x; q < ex.e_off[e + 1]; q += blockDim.x) {
    const unsigned code = ex.e_ij[q];
    const long long i = code & 0xffffu, j = code >> 16;
    cfma(acc, ex.e_val[q], Y[j * NP + i]);      // E_ij rho_ji
  }
  acc = block_sum_c(acc, red);
  if (threadIdx.x == 0) {
    const long long off = ex.base + (long long)e * ex.n_out + dev->emit_o;
    if (ex.complex_out) ((cplx*)ex.expect)[off] = acc; else ((double*)ex.expect)[off] = acc.x;
  }
}
__global__ void lh_expect_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, const __grid_constant__ LargeHermExpect ex) {
  __shared__ double red[34];
  lh_expect_body(dev, blockIdx.x, ex, bf.Yfull[0], bf.NPh, red);
}

// full matrix 0 -> interleaved [N][N] (stored states: only when the committed state is an output; dev == nullptr: always)
__global__ void lh_unpack_kernel(const LargeDev* __restrict__ dev, const cplx* __restrict__ src, cplx* __restrict__ dst, long long per_output, int N, int NP) {
  if (dev) {
    if (!dev->emit) return;
    dst += (long long)dev->emit_o * per_output;
  }
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / N;
    const int j = (int)(idx - i * N);
    dst[idx] = src[i * NP + j];
  }
}

// input of the attempt's first stage, y + h a_10 k1 -> full matrix 1.  (st0 = 0 - k1 not known yet, the first attempt of a
// trajectory - needs nothing: the stage input is y itself, which lh_pack_kernel left in matrix 0.)
// Blocks behind the tiles: expectation values of the committed state (matrix 0, which this kernel does not write), so
// that an output costs no launch of its own.
__global__ void __launch_bounds__(256) lh_first_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int n_tiles,
                                                       const __grid_constant__ LargeHermExpect ex) {
  __shared__ cplx sm[LH_T * (LH_T + 1)];
  lh_grid_dependency_wait();
  if ((int)blockIdx.x >= n_tiles) {
    lh_expect_body(dev, (int)blockIdx.x - n_tiles, ex, bf.Yfull[0], bf.NPh, (double*)sm);
    return;
  }
  if (!dev->active) return;
  if (dev->st0 == 0) return;
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long tbase = (long long)blockIdx.x * LH_TE;
  const cplx* y = bf.ybuf[dev->ycur] + tbase;
  const cplx* k0 = bf.kbuf[dev->kmap[0]] + tbase;
  const double h = dev->htry, a10 = RK_A[dev->ctl.pair][1][0];
  cplx v[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int idx = (w + 8 * q) * LH_T + lane;
    const cplx yv = LQ_LD(y + idx), kv = LQ_LD(k0 + idx);
    v[q] = cmake(fma(h, fma(a10, kv.x, 0.0), yv.x), fma(h, fma(a10, kv.y, 0.0), yv.y));
  }
  lh_write_tile(bf.Yfull[1], bf.NPh, ti, tj, v, sm);
}

// stage st of the attempt on one upper tile: right-hand side, then the next stage input (or, last stage, the error sum)
template <int MINB, int UNR>
__global__ void __launch_bounds__(256, MINB) lh_stage_kernel(LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int st,
                                                          const __grid_constant__ LargeDiagArgs a, double atol, double rtol) {
  extern __shared__ __align__(16) unsigned char lh_smem[];
  lh_grid_dependency_wait();
  if (!dev->active || st < dev->st0 || st >= dev->S) return;
  cplx* sm = (cplx*)lh_smem;                               // transpose buffer 32 x 33
  cplx* gts = sm + LH_T * (LH_T + 1);                      // [nD][rows 32 | columns 32]: G(t) diagonals of this stage time
  cplx* lams = gts + bf.nD * 2 * LH_T;                     // [nT][rows | columns]: collapse diagonals
  cplx* gds = lams + bf.nT * 2 * LH_T;                     // [rows | columns]: static main diagonal of G0
  double* red = (double*)(gds + 2 * LH_T);
  const int pair = dev->ctl.pair, S = dev->S, N = bf.N, NP = bf.NPh;
  const bool last = st == S - 1;
  const double h = dev->htry;
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long tbase = (long long)blockIdx.x * LH_TE;
  const double* cval = dev->cval + st * LDEV_MAXG;

  // ---- the epilogue's packed streams (y and the earlier stage derivatives of this tile, 16 KB each) start their way from
  // HBM into L2 now, behind the gathers: one 128-byte line per thread and vector
  const cplx* y = bf.ybuf[dev->ycur] + tbase;
  if (threadIdx.x < LH_TE * (int)sizeof(cplx) / 128) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)y + threadIdx.x * 128));
    for (int jj = 0; jj < st; jj++) {
      if ((last ? RK_E[pair][jj] : RK_A[pair][st + 1][jj]) == 0.0) continue;
      asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(bf.kbuf[dev->kmap[jj]] + tbase) + threadIdx.x * 128));
    }
  }
  // ---- tile tables (entries beyond N are zero: padded elements then get zero weights and come out as exact zeros)
  for (int q = threadIdx.x; q < (bf.nD + bf.nT + 1) * 2 * LH_T; q += blockDim.x) {
    const int d = q / (2 * LH_T), r = q % LH_T;
    const int pos = (((q / LH_T) & 1) ? tj : ti) * LH_T + r;
    cplx g = cmake(0, 0);
    if (pos < N && !(bf.dbg & 32)) {
      if (d < bf.nD) {      // g0 + sum_m c_m(t) gk_m
        g = bf.g0[(long long)d * N + pos];
        for (int m = 0; m < bf.n_g; m++) {
          const cplx gm = bf.gk[((long long)m * bf.nD + d) * N + pos];
          const double c = cval[m];
          g.x = fma(c, gm.x, g.x);
          g.y = fma(c, gm.y, g.y);
        }
      } else if (d < bf.nD + bf.nT) {
        g = bf.lam[(long long)(d - bf.nD) * N + pos];
      } else {
        g = bf.gdiag[pos];
      }
    }
    gts[q] = g;
  }
  __syncthreads();

  // ---- right-hand side of 4 elements: rows w, w + 8, w + 16, w + 24 of column `lane`
  const cplx* __restrict__ Y = bf.Yfull[st & 1];
  const cplx* p0 = Y + (long long)(ti * LH_T + w) * NP + tj * LH_T + lane;
  const long long rs = 8LL * NP;
  const int cj = LH_T + lane;      // column entry of a tile table
  cplx kv[4];
  {
    cplx yv[4], wt[4];
    const cplx gj = gds[cj];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      yv[q] = p0[q * rs];
      const cplx gi = gds[w + 8 * q];
      wt[q] = cmake(gi.y + gj.y, -(gi.x - gj.x));      // -i (G0_ii - conj(G0_jj))
    }
    for (int t = 0; t < a.nLs; t++) {
      const cplx lb = lams[a.ls_b[t] * 64 + cj];
      const cplx* la_ = lams + a.ls_a[t] * 64 + w;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const cplx la = la_[8 * q];
        wt[q].x += la.x * lb.x + la.y * lb.y;
        wt[q].y += la.y * lb.x - la.x * lb.y;
      }
    }
    if (a.g0_tab >= 0) {
      const cplx vj = gts[a.g0_tab * 64 + cj];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const cplx vi = gts[a.g0_tab * 64 + w + 8 * q];
        wt[q].x += vi.y + vj.y;
        wt[q].y -= vi.x - vj.x;
      }
    }
    for (int t = 0; t < a.nLo; t++) {
      const double rr = cval[a.lo_r[t]];
      const cplx lb = lams[a.lo_b[t] * 64 + cj];
      const cplx* la_ = lams + a.lo_a[t] * 64 + w;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const cplx la = la_[8 * q];
        wt[q].x += rr * (la.x * lb.x + la.y * lb.y);
        wt[q].y += rr * (la.y * lb.x - la.x * lb.y);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) kv[q] = cmul(wt[q], yv[q]);
  }
#pragma unroll UNR
  for (int t = 0; t < ((bf.dbg & 64) ? 0 : a.nGo); t++) {
    const int o = a.go_off[t];
    const cplx* gv = gts + a.go_tab[t] * 64;
    const cplx vj = gv[cj];
    const cplx cjv = cmake(vj.y, vj.x);                        // +i conj(v_d[j])
    const cplx* pr = p0 + (long long)o * NP;                   // Y[i + o, j]
    const cplx* pc = p0 + o;                                   // Y[i, j + o]
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const cplx vi = gv[w + 8 * q];
      const cplx ya = pr[q * rs], yb = pc[q * rs];
      cfma(kv[q], cmake(vi.y, -vi.x), ya);                     // -i v_d[i] Y[i+o, j]
      cfma(kv[q], cjv, yb);
    }
  }
  for (int t = 0; t < ((bf.dbg & 64) ? 0 : a.nLg); t++) {
    const int ri = a.lg_r[t];
    const double rr = ri < 0 ? 1.0 : cval[ri];
    const cplx lb = lams[a.lg_b[t] * 64 + cj];
    const cplx* la_ = lams + a.lg_a[t] * 64 + w;
    const cplx* pg = p0 + (long long)a.lg_oa[t] * NP + a.lg_ob[t];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const cplx la = la_[8 * q];
      const cplx wv = cmake(rr * (la.x * lb.x + la.y * lb.y), rr * (la.y * lb.x - la.x * lb.y));
      cfma(kv[q], wv, pg[q * rs]);
    }
  }
  {
    cplx* kd = bf.kbuf[dev->kmap[st]] + tbase;
    if (!(bf.dbg & 4))
#pragma unroll
    for (int q = 0; q < 4; q++) LQ_ST(kd + (w + 8 * q) * LH_T + lane, kv[q]);
  }
  if (!last) {
    // Y_{st+1} = y + h sum_{jj <= st} a_{st+1,jj} k_jj   (zero coefficients skipped: fma(0, k, acc) = acc)
    cplx acc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) acc[q] = cmake(0, 0);
    for (int jj = 0; jj < ((bf.dbg & 128) ? 0 : st); jj++) {
      const double aj = RK_A[pair][st + 1][jj];
      if (aj == 0.0) continue;
      const cplx* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const cplx kq = LQ_LD(kj + (w + 8 * q) * LH_T + lane);
        acc[q].x = fma(aj, kq.x, acc[q].x);
        acc[q].y = fma(aj, kq.y, acc[q].y);
      }
    }
    const double as = RK_A[pair][st + 1][st];
    cplx v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const cplx yq = LQ_LD(y + (w + 8 * q) * LH_T + lane);
      acc[q].x = fma(as, kv[q].x, acc[q].x);
      acc[q].y = fma(as, kv[q].y, acc[q].y);
      v[q] = cmake(fma(h, acc[q].x, yq.x), fma(h, acc[q].y, yq.y));
    }
    if (st == S - 2) {      // the propagated solution: the new state if the step is accepted
      cplx* yn = bf.ybuf[dev->ycur ^ 1] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) LQ_ST(yn + (w + 8 * q) * LH_T + lane, v[q]);
    }
    if (bf.dbg & 3) {
      if (!(bf.dbg & 2)) lh_write_tile(bf.Yfull[(st + 1) & 1], NP, ti, ti, v, sm);      // own tile only (written where the diagonal tile would go: timing only)
    } else
    lh_write_tile(bf.Yfull[(st + 1) & 1], NP, ti, tj, v, sm);
  } else {
    // error estimate h sum_jj e_jj k_jj, weighted by atol + rtol max(|y|, |y_new|); the stage input of the last stage
    // is y_new
    cplx er[4];
#pragma unroll
    for (int q = 0; q < 4; q++) er[q] = cmake(0, 0);
    for (int jj = 0; jj < S - 1; jj++) {
      const double ej = RK_E[pair][jj];
      if (ej == 0.0) continue;
      const cplx* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const cplx kq = LQ_LD(kj + (w + 8 * q) * LH_T + lane);
        er[q].x = fma(ej, kq.x, er[q].x);
        er[q].y = fma(ej, kq.y, er[q].y);
      }
    }
    // (most of rho is exact zeros: the IEEE square root of an exact zero and a division with a zero numerator take their
    // slow paths - a floor far below atol under the root and a reciprocal of the never-zero weight avoid both)
    const double el = RK_E[pair][S - 1];
    const double m2floor = atol >= 1e-100 ? 1e-280 : 0.0;
    double es = 0.0;
    const int j = tj * LH_T + lane;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int r = w + 8 * q;
      if (ti * LH_T + r < N && j < N) {
        const double ex = fma(el, kv[q].x, er[q].x) * h, ey = fma(el, kv[q].y, er[q].y) * h;
        const cplx yq = LQ_LD(y + r * LH_T + lane), nq = p0[q * rs];
        const double s = atol + rtol * sqrt(fmax(fmax(yq.x * yq.x + yq.y * yq.y, nq.x * nq.x + nq.y * nq.y), m2floor));
        es = fma(ex * ex + ey * ey, 1.0 / (s * s), es);
      }
    }
    es = block_sum(es, red);
    if (threadIdx.x == 0) atomicAdd(&dev->err2, ti == tj ? es : 2.0 * es);
  }
}

}  // namespace seq
