// Large path (C5), Hermitian tiles: the device-controlled stepper of kernel_large_dev.cuh on HALF of rho, with the stage
// combination and the error estimate fused into the right-hand-side kernel.
//
// rho is Hermitian and so is every stage derivative (the Lindblad right-hand side maps Hermitian matrices to Hermitian
// matrices), so only the 32 x 32 tiles (ti <= tj) of the upper block triangle are integrated.  State and stage
// derivatives live PACKED ([tile][re | im][32 x 32], half the bytes of a full vector); the stage input alone is kept
// as a full planar matrix (two of them, ping-pong), because the structured right-hand side gathers Y[i + o, j + o']
// from anywhere.  One kernel per stage and tile does
//     k_st   = f(t_st, Y_st)                        gathers from the full stage matrix (L1 / L2)
//     Y_st+1 = y + h sum_{j <= st} a_{st+1,j} k_j    packed streams (HBM), k_st from registers
// and writes Y_st+1 into the other full matrix: its own tile and - through a shared-memory transpose, conjugated - the
// mirrored tile.  The last stage's epilogue is the error estimate instead (weight 2 off the block diagonal).  Per 4(3)
// step that is 4 fused kernels + 1 stage-input kernel over ~36 half vectors instead of 4 + 4 + 1 kernels over ~68:
// the arithmetic of an element (order of the multiply-adds) is the one of ldev_combine_kernel / ldev_err_kernel.
// Stage st reads full matrix st & 1 and writes (st + 1) & 1; both pairs have an even last stage, so the committed
// state always ends in matrix 0, where the output kernels (ldev_expect_kernel / ldev_store_kernel) read it.
#pragma once
#include "kernel_large_dev.cuh"

namespace seq {

#define LH_T 32
#define LH_TE (LH_T * LH_T)

struct LargeHermBufs {
  double* ybuf[2];          // packed state / new state
  double* kbuf[7];          // packed stage derivatives
  double* Yfull[2];         // full stage matrices, planar [2][NPh][NPh]
  const unsigned* tiles;    // ti | tj << 16 of every upper tile
  long long yplane;         // NPh * NPh
  int NPh, N;
  const cplx *g0, *gk;      // G diagonals: static part [nD][N] and time-dependent terms [n_g][nD][N] (diag_fill_kernel's tables)
  int n_g, nD;
};

// v[q]: element (row w + 8 q, column lane) of tile (ti, tj) -> the tile and (ti != tj) its conjugate-transposed mirror
__device__ __forceinline__ void lh_write_tile(double* __restrict__ Yo, long long yplane, int NP, int ti, int tj, const cplx (&v)[4],
                                              double (&sm)[2][LH_T][LH_T + 1]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int r = w + 8 * q;
    const long long o = (long long)(ti * LH_T + r) * NP + tj * LH_T + lane;
    Yo[o] = v[q].x;
    Yo[yplane + o] = v[q].y;
  }
  if (ti != tj) {      // (block-uniform)
#pragma unroll
    for (int q = 0; q < 4; q++) { sm[0][w + 8 * q][lane] = v[q].x; sm[1][w + 8 * q][lane] = v[q].y; }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int r = w + 8 * q;
      const long long o = (long long)(tj * LH_T + r) * NP + ti * LH_T + lane;
      Yo[o] = sm[0][lane][r];
      Yo[yplane + o] = -sm[1][lane][r];
    }
  }
}

// rho0 (interleaved [N][N]) -> packed state + full matrix 0; chk[0] = max |rho_ij|^2, chk[1] = max |rho_ji - conj(rho_ij)|^2
// as bit patterns of non-negative doubles (the host falls back to the full-matrix stepper for a non-Hermitian rho0)
__global__ void __launch_bounds__(256) lh_pack_kernel(const cplx* __restrict__ src, const __grid_constant__ LargeHermBufs bf, unsigned long long* __restrict__ chk) {
  __shared__ double sm[2][LH_T][LH_T + 1];
  __shared__ double red[34];
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16, N = bf.N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* yp = bf.ybuf[0] + (long long)blockIdx.x * (2 * LH_TE);
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
    yp[r * LH_T + lane] = v[q].x;
    yp[LH_TE + r * LH_T + lane] = v[q].y;
  }
  lh_write_tile(bf.Yfull[0], bf.yplane, bf.NPh, ti, tj, v, sm);
  // block maxima (values are >= 0: the integer order of the bit patterns is the order of the numbers)
  for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); as = fmax(as, __shfl_xor_sync(0xffffffffu, as, o)); }
  __syncthreads();
  if (lane == 0) { red[w] = mx; red[8 + w] = as; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) { mx = fmax(mx, red[q]); as = fmax(as, red[8 + q]); }
    atomicMax(&chk[0], (unsigned long long)__double_as_longlong(mx));
    atomicMax(&chk[1], (unsigned long long)__double_as_longlong(as));
  }
}

// input of the attempt's first stage, y + h a_10 k1 -> full matrix 1.  (st0 = 0 - k1 not known yet, the first attempt of a
// trajectory - needs nothing: the stage input is y itself, which lh_pack_kernel left in matrix 0.)
// Blocks behind the tiles: expectation values of the committed state (matrix 0, which this kernel does not write), so
// that an output costs no launch of its own.
__global__ void __launch_bounds__(256) lh_first_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int n_tiles,
                                                       const __grid_constant__ LargeExpectArgs ex) {
  __shared__ double sm[2][LH_T][LH_T + 1];
  if ((int)blockIdx.x >= n_tiles) {
    ldev_expect_body(dev, (int)blockIdx.x - n_tiles, ex.e_off, ex.e_ij, ex.e_val, ex.Y, ex.y_plane, ex.NP, ex.expect, ex.base, ex.n_out, ex.complex_out, &sm[0][0][0]);
    return;
  }
  if (!dev->active) return;
  const int st0 = dev->st0;
  if (st0 == 0) return;
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long tbase = (long long)blockIdx.x * (2 * LH_TE);
  const double* y = bf.ybuf[dev->ycur] + tbase;
  const double* k0 = bf.kbuf[dev->kmap[0]] + tbase;
  const double h = dev->htry, a10 = RK_A[dev->ctl.pair][1][0];
  cplx v[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int idx = (w + 8 * q) * LH_T + lane;
    v[q] = cmake(fma(h, fma(a10, k0[idx], 0.0), y[idx]), fma(h, fma(a10, k0[LH_TE + idx], 0.0), y[LH_TE + idx]));
  }
  lh_write_tile(bf.Yfull[1], bf.yplane, bf.NPh, ti, tj, v, sm);
}

// stage st of the attempt on one upper tile: right-hand side, then the next stage input (or, last stage, the error sum)
__global__ void __launch_bounds__(256) lh_stage_kernel(LargeDev* __restrict__ dev, const __grid_constant__ LargeHermBufs bf, int st,
                                                       const __grid_constant__ LargeDiagArgs a, double atol, double rtol) {
  __shared__ double sm[2][LH_T][LH_T + 1];
  __shared__ double red[34];
  __shared__ cplx gts[DIAG_MAX_GD * 2 * LH_T];
  if (!dev->active || st < dev->st0 || st >= dev->S) return;
  const int pair = dev->ctl.pair, S = dev->S, N = bf.N, NP = bf.NPh;
  const bool last = st == S - 1;
  const double h = dev->htry;
  const unsigned code = bf.tiles[blockIdx.x];
  const int ti = code & 0xffffu, tj = code >> 16;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long tbase = (long long)blockIdx.x * (2 * LH_TE);
  const double* Yr = bf.Yfull[st & 1];
  const double* Yi = Yr + bf.yplane;
  const double* cval = dev->cval + st * LDEV_MAXG;
  const int j = tj * LH_T + lane;

  // G(t) diagonals of this stage time at the tile's 32 rows and 32 columns: g0 + sum_m c_m(t) gk_m
  for (int q = threadIdx.x; q < bf.nD * 2 * LH_T; q += blockDim.x) {
    const int d = q / (2 * LH_T), r = q % LH_T;
    const int pos = (((q / LH_T) & 1) ? tj : ti) * LH_T + r;
    cplx g = cmake(0, 0);
    if (pos < N) {
      g = bf.g0[(long long)d * N + pos];
      for (int m = 0; m < bf.n_g; m++) {
        const cplx gm = bf.gk[((long long)m * bf.nD + d) * N + pos];
        const double c = cval[m];
        g.x = fma(c, gm.x, g.x);
        g.y = fma(c, gm.y, g.y);
      }
    }
    gts[q] = g;
  }
  __syncthreads();

  cplx kv[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int r = w + 8 * q, i = ti * LH_T + r;
    kv[q] = (i < N && j < N) ? large_diag_rhs_elem(a, Yr, Yi, NP, cval, i, j, LargeGtTile{gts, r, lane}) : cmake(0, 0);
  }
  {
    double* kd = bf.kbuf[dev->kmap[st]] + tbase;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int idx = (w + 8 * q) * LH_T + lane;
      kd[idx] = kv[q].x;
      kd[LH_TE + idx] = kv[q].y;
    }
  }
  const double* y = bf.ybuf[dev->ycur] + tbase;
  if (!last) {
    // Y_{st+1} = y + h sum_{jj <= st} a_{st+1,jj} k_jj   (zero coefficients skipped: fma(0, k, acc) = acc)
    cplx acc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) acc[q] = cmake(0, 0);
    for (int jj = 0; jj < st; jj++) {
      const double aj = RK_A[pair][st + 1][jj];
      if (aj == 0.0) continue;
      const double* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int idx = (w + 8 * q) * LH_T + lane;
        acc[q].x = fma(aj, kj[idx], acc[q].x);
        acc[q].y = fma(aj, kj[LH_TE + idx], acc[q].y);
      }
    }
    const double as = RK_A[pair][st + 1][st];
    cplx v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int idx = (w + 8 * q) * LH_T + lane;
      acc[q].x = fma(as, kv[q].x, acc[q].x);
      acc[q].y = fma(as, kv[q].y, acc[q].y);
      v[q] = cmake(fma(h, acc[q].x, y[idx]), fma(h, acc[q].y, y[LH_TE + idx]));
    }
    if (st == S - 2) {      // the propagated solution: the new state if the step is accepted
      double* yn = bf.ybuf[dev->ycur ^ 1] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int idx = (w + 8 * q) * LH_T + lane;
        yn[idx] = v[q].x;
        yn[LH_TE + idx] = v[q].y;
      }
    }
    lh_write_tile(bf.Yfull[(st + 1) & 1], bf.yplane, NP, ti, tj, v, sm);
  } else {
    // error estimate h sum_jj e_jj k_jj, weighted by atol + rtol max(|y|, |y_new|); the stage input of the last stage
    // is y_new
    cplx er[4];
#pragma unroll
    for (int q = 0; q < 4; q++) er[q] = cmake(0, 0);
    for (int jj = 0; jj < S - 1; jj++) {
      const double ej = RK_E[pair][jj];
      if (ej == 0.0) continue;
      const double* kj = bf.kbuf[dev->kmap[jj]] + tbase;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int idx = (w + 8 * q) * LH_T + lane;
        er[q].x = fma(ej, kj[idx], er[q].x);
        er[q].y = fma(ej, kj[LH_TE + idx], er[q].y);
      }
    }
    const double el = RK_E[pair][S - 1];
    double es = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int r = w + 8 * q, idx = r * LH_T + lane;
      const long long i = ti * LH_T + r;
      if (i < N && j < N) {
        const double ex = fma(el, kv[q].x, er[q].x) * h, ey = fma(el, kv[q].y, er[q].y) * h;
        const double yr = y[idx], yi = y[LH_TE + idx];
        const double nr = Yr[i * NP + j], ni = Yi[i * NP + j];
        const double s = atol + rtol * sqrt(fmax(yr * yr + yi * yi, nr * nr + ni * ni));
        es += (ex * ex + ey * ey) / (s * s);
      }
    }
    es = block_sum(es, red);
    if (threadIdx.x == 0) atomicAdd(&dev->err2, ti == tj ? es : 2.0 * es);
  }
}

}  // namespace seq
