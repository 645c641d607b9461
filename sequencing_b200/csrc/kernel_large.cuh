// SEQ_METHOD_LARGE: one big Lindblad system (N > 96, C5: N = 1875) whose rho does not fit in one
// SM.  The whole GPU (or, row-sharded, every GPU of the box) works on ONE right-hand side:
//
//   K[R,:] = -i G(t)[R,:] Y + i Y[R,:] G(t)^dag + sum_l r_l (L_l[R,:] Y) L_l^dag          (R = this rank's rows)
//
// i.e. 2 + 2 n_l complex (m x NP x NP) products per evaluation, each run by zgemm_dmma_kernel:
//  * FP64 tensor cores (DMMA.8x8x4), complex product in Karatsuba form (3 real chains instead of 4);
//  * operands PLANAR (re plane, im plane), row major, zero padded to NP = 16*ceil(N/16), so that
//    global -> shared copies are 16-byte cp.async streams (3-stage pipeline, K chunks of 16);
//  * 48 x 96 CTA tiles, 12 warps each owning 24 x 16 (3 x 2 DMMA tiles);
//  * STREAM-K scheduling: one CTA per SM, the (tile, K-chunk) iteration space is cut into equal
//    contiguous ranges, so the SMs stay evenly loaded whatever the tile count is (row-sharded C5 has
//    5 x 20 = 100 tiles for 148 SMs).  Partial tiles are combined with FP64 reductions (red.add.f64)
//    into a pre-initialised C, which is also how the products accumulate into K.
//  * operator CHUNK SPARSITY: every product has an operator (G or L_l) as its A operand (NN) or as its
//    conj-transposed B operand (NH).  large_sched_kernel lists, per operator row tile, the K chunks
//    (16 columns) that hold a non-zero; the stream-K iteration space is built from those lists only, so
//    Kronecker-structured operators (C5: ~4 of 118 chunks per tile) cost what they contain.  Dense
//    operators list every chunk.
// The Runge-Kutta driver (host loop, one device sync per step for the error norm) and the
// row-sharding exchange (all-gather of the stage state per RHS) live in seq_engine.cu.
#pragma once
#include "kernel_dmma.cuh"

namespace seq {

struct LargeGemmCfg {
  static constexpr int TM = 48, TN = 96, KC = 16, NST = 3;
  static constexpr int NW = 12, NTHREADS = NW * 32;
  static constexpr int LDA = KC + 4;        // A tile / B^H tile rows: [row][KC] (+4: conflict-free fragments)
  static constexpr int LDB = TN + 4;        // B tile (NN): [KC][TN]
  static constexpr int A_PLANE = TM * LDA;  // doubles
  static constexpr int B_PLANE = (TN * LDA > KC * LDB) ? TN * LDA : KC * LDB;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr size_t SMEM = (size_t)NST * STAGE * sizeof(double);
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

struct ZgemmArgs {
  const double* Ar; const double* Ai; long long lda;   // A: M x K
  const double* Br; const double* Bi; long long ldb;   // BH ? (Ncols x K, used as conj-transpose) : (K x Ncols)
  double* Cr; double* Ci; long long ldc;               // C: M x Ncols, pre-initialised; C += alpha * scale * A op(B)
  int M, Ncols, K;                                     // K multiple of 16; Ncols even
  double alpha_r, alpha_i;
  const double* scale;                                 // optional device scalar (time-dependent collapse rate)
  // chunk schedule of the operator operand (A for NN: per 48-row tile; B for NH: per 96-row tile):
  // pref[t] = number of listed chunks in tiles < t (pref[tiles] = all), idx[t * kchunks + i] = i-th listed chunk of tile t
  const int* pref;
  const unsigned short* idx;
};

template <bool BH>
__global__ void __launch_bounds__(LargeGemmCfg::NTHREADS, 1) zgemm_dmma_kernel(const __grid_constant__ ZgemmArgs a) {
  typedef LargeGemmCfg C;
  extern __shared__ __align__(16) double smem_large[];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int wr = w / 6, wc = w % 6;   // warp grid 2 x 6 -> rows 24*wr, cols 16*wc
  const int tiles_m = (a.M + C::TM - 1) / C::TM, tiles_n = (a.Ncols + C::TN - 1) / C::TN;
  const int kchunks = a.K / C::KC;
  // iteration space: tiles ordered tm-fastest; a tile owns the listed chunks of its operator row tile
  // (NN: row tile tm of A; NH: row tile tn of B)
  const int nsched = BH ? tiles_n : tiles_m;
  const long long listed = a.pref[nsched];
  const long long total = listed * (BH ? tiles_m : tiles_n);
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  long long it = per * blockIdx.x;
  const long long it_end = (it + per < total) ? it + per : total;
  const double sc = a.scale ? *a.scale : 1.0;
  const double alr = a.alpha_r * sc, ali = a.alpha_i * sc;

  while (it < it_end) {
    int tm, tn, pos0, cnt;
    if (BH) {
      // per tn: tiles_m * cnt[tn] iterations
      tn = 0;
      while (tn + 1 < tiles_n && (long long)a.pref[tn + 1] * tiles_m <= it) tn++;
      cnt = a.pref[tn + 1] - a.pref[tn];
      const long long r = it - (long long)a.pref[tn] * tiles_m;
      tm = (int)(r / cnt);
      pos0 = (int)(r - (long long)tm * cnt);
    } else {
      // per tn: listed iterations, split over tm by pref
      tn = (int)(it / listed);
      const int r = (int)(it - (long long)tn * listed);
      tm = 0;
      while (tm + 1 < tiles_m && a.pref[tm + 1] <= r) tm++;
      cnt = a.pref[tm + 1] - a.pref[tm];
      pos0 = r - a.pref[tm];
    }
    const long long left = it_end - it;
    const int pos1 = (int)((cnt - pos0 < left) ? cnt : pos0 + left);
    const unsigned short* list = a.idx + (long long)(BH ? tn : tm) * kchunks;
    const int kc0 = pos0, kc1 = pos1;    // positions in the chunk list
    const int m0 = tm * C::TM, n0 = tn * C::TN;

    auto load_stage = [&](int stage, int kc) {
      double* As = smem_large + stage * C::STAGE;
      double* Bs = As + 2 * C::A_PLANE;
      const int k0 = (int)list[kc] * C::KC;
      // A: 2 planes x 48 rows x 8 chunks
      for (int c = tid; c < 2 * C::TM * (C::KC / 2); c += C::NTHREADS) {
        const int pl = c / (C::TM * (C::KC / 2)), rem = c - pl * (C::TM * (C::KC / 2));
        const int r = rem / (C::KC / 2), ch = rem - r * (C::KC / 2);
        const bool ok = (m0 + r) < a.M;
        const double* src = (pl ? a.Ai : a.Ar) + (long long)(ok ? m0 + r : 0) * a.lda + k0 + 2 * ch;
        cp_async16_zfill(As + pl * C::A_PLANE + r * C::LDA + 2 * ch, src, ok);
      }
      if (BH) {
        // B^H operand: rows n of B, K contiguous: 2 planes x 96 rows x 8 chunks
        for (int c = tid; c < 2 * C::TN * (C::KC / 2); c += C::NTHREADS) {
          const int pl = c / (C::TN * (C::KC / 2)), rem = c - pl * (C::TN * (C::KC / 2));
          const int r = rem / (C::KC / 2), ch = rem - r * (C::KC / 2);
          const bool ok = (n0 + r) < a.Ncols;
          const double* src = (pl ? a.Bi : a.Br) + (long long)(ok ? n0 + r : 0) * a.ldb + k0 + 2 * ch;
          cp_async16_zfill(Bs + pl * C::B_PLANE + r * C::LDA + 2 * ch, src, ok);
        }
      } else {
        // B operand: 2 planes x 16 rows (k) x 48 chunks
        for (int c = tid; c < 2 * C::KC * (C::TN / 2); c += C::NTHREADS) {
          const int pl = c / (C::KC * (C::TN / 2)), rem = c - pl * (C::KC * (C::TN / 2));
          const int r = rem / (C::TN / 2), ch = rem - r * (C::TN / 2);
          const bool ok = (n0 + 2 * ch) < a.Ncols;
          const double* src = (pl ? a.Bi : a.Br) + (long long)(k0 + r) * a.ldb + (ok ? n0 + 2 * ch : 0);
          cp_async16_zfill(Bs + pl * C::B_PLANE + r * C::LDB + 2 * ch, src, ok);
        }
      }
    };

    double P[3][2][3][2];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int q = 0; q < 3; q++) { P[i][j][q][0] = 0.0; P[i][j][q][1] = 0.0; }

    __syncthreads();   // every warp is done with the stages of the previous segment
#pragma unroll
    for (int s = 0; s < C::NST - 1; s++) {
      if (kc0 + s < kc1) load_stage(s, kc0 + s);
      cp_async_commit();
    }
    for (int kc = kc0; kc < kc1; kc++) {
      cp_async_wait<C::NST - 2>();
      __syncthreads();
      {
        const int nk = kc + C::NST - 1;
        if (nk < kc1) load_stage((nk - kc0) % C::NST, nk);
        cp_async_commit();
      }
      const double* As = smem_large + ((kc - kc0) % C::NST) * C::STAGE;
      const double* Bs = As + 2 * C::A_PLANE;
#pragma unroll
      for (int ks = 0; ks < C::KC / 4; ks++) {
        const int k0 = 4 * ks;
        double ar[3], ai[3], as[3], br[2], bi[2], bs[2];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const double* pa = As + (24 * wr + 8 * i + g) * C::LDA + k0 + tg;
          ar[i] = pa[0]; ai[i] = pa[C::A_PLANE]; as[i] = ar[i] + ai[i];
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          if (BH) {
            const double* pb = Bs + (16 * wc + 8 * j + g) * C::LDA + k0 + tg;
            br[j] = pb[0]; bi[j] = pb[C::B_PLANE]; bs[j] = br[j] - bi[j];
          } else {
            const double* pb = Bs + (k0 + tg) * C::LDB + 16 * wc + 8 * j + g;
            br[j] = pb[0]; bi[j] = pb[C::B_PLANE]; bs[j] = br[j] + bi[j];
          }
        }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) {
            dmma884(P[i][j][0][0], P[i][j][0][1], ar[i], br[j]);
            dmma884(P[i][j][1][0], P[i][j][1][1], ai[i], bi[j]);
            dmma884(P[i][j][2][0], P[i][j][2][1], as[i], bs[j]);
          }
      }
    }
    cp_async_wait<0>();
    // epilogue: C += alpha * (A op(B)) for this K range (FP64 reductions; other CTAs may hold the rest of K)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int row = m0 + 24 * wr + 8 * i + g;
        const int col = n0 + 16 * wc + 8 * j + 2 * tg;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double cr, ci;
          if (BH) { cr = P[i][j][0][c] + P[i][j][1][c]; ci = P[i][j][2][c] - P[i][j][0][c] + P[i][j][1][c]; }
          else { cr = P[i][j][0][c] - P[i][j][1][c]; ci = P[i][j][2][c] - P[i][j][0][c] - P[i][j][1][c]; }
          if (row < a.M && col + c < a.Ncols) {
            const long long off = (long long)row * a.ldc + col + c;
            atomicAdd(a.Cr + off, alr * cr - ali * ci);
            atomicAdd(a.Ci + off, alr * ci + ali * cr);
          }
        }
      }
    it += (kc1 - kc0);
  }
}

// ---- elementwise kernels of the large path (all planar, leading dimension NP) -----------------
// interleaved complex [N,N] (row major) -> planar padded [2][rows_pad][NP] (zero padded)
__global__ void large_pack_kernel(const cplx* __restrict__ src, double* __restrict__ dst, int N, int NP, long long rows_pad) {
  const long long plane = rows_pad * NP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / NP;
    const int j = (int)(idx - i * NP);
    cplx v = cmake(0, 0);
    if (i < N && j < N) v = src[i * N + j];
    dst[idx] = v.x;
    dst[plane + idx] = v.y;
  }
}
// planar padded -> interleaved complex [N,N]
__global__ void large_unpack_kernel(const double* __restrict__ src, cplx* __restrict__ dst, int N, int NP, long long rows_pad) {
  const long long plane = rows_pad * NP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / N;
    const int j = (int)(idx - i * N);
    dst[idx] = cmake(src[i * NP + j], src[plane + i * NP + j]);
  }
}
// spline values of every time-dependent term at time t -> cval[n_g]
__global__ void large_coeff_kernel(SolveParams p, int b, double t, double* __restrict__ cval) {
  const int m = threadIdx.x;
  if (m >= p.n_g) return;
  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  double v;
  if (m < p.n_ch) {
    v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, t);
    if (p.coeff_scale) v *= p.coeff_scale[(long long)b * p.n_ch + m];
  } else {
    v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, t);
  }
  cval[m] = v;
}
// G(t) = G0 + sum_m cval[m] Gk[m] over both planes (n doubles per matrix)
__global__ void large_build_G_kernel(const double* __restrict__ G0, const double* __restrict__ Gk, const double* __restrict__ cval,
                                     int n_g, long long n, double* __restrict__ G) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = G0[i];
    for (int m = 0; m < n_g; m++) v = fma(cval[m], Gk[(long long)m * n + i], v);
    G[i] = v;
  }
}
// dst = y + h sum_j a[j] k_j  over this rank's rows (both planes; `n` doubles per plane, plane strides differ:
// the k/y buffers are [2][m][NP], dst is the rank's row block inside the full [2][rows_full][NP] stage buffer)
struct LargeCombine {
  const double* y; const double* k[7]; double a[7]; int nk; double h;
  long long n_plane;          // m * NP
  double* dst; long long dst_plane;   // plane stride of dst
  double* keep;               // optional copy in local layout (y_new)
};
__global__ void large_combine_kernel(const __grid_constant__ LargeCombine c) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * c.n_plane; i += (long long)gridDim.x * blockDim.x) {
    const int pl = i >= c.n_plane;
    const long long r = i - pl * c.n_plane;
    double acc = 0.0;
    for (int j = 0; j < c.nk; j++) acc = fma(c.a[j], c.k[j][i], acc);
    const double v = fma(c.h, acc, c.y[i]);
    c.dst[pl * c.dst_plane + r] = v;
    if (c.keep) c.keep[i] = v;
  }
}
// err2 += sum over the physical elements of this rank's rows of |h sum_j e_j k_j|^2 / (atol + rtol max(|y|,|ynew|))^2
struct LargeErr {
  const double* y; const double* yn; const double* k[7]; double e[7]; int nk; double h, atol, rtol;
  int m, NP, N; long long row0;
  double* out;
};
__global__ void large_err_kernel(const __grid_constant__ LargeErr c) {
  __shared__ double red[34];
  const long long n_plane = (long long)c.m * c.NP;
  double es = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_plane; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c.NP;
    const int col = (int)(i - r * c.NP);
    if (c.row0 + r >= c.N || col >= c.N) continue;
    double er = 0.0, ei = 0.0;
    for (int j = 0; j < c.nk; j++) { er = fma(c.e[j], c.k[j][i], er); ei = fma(c.e[j], c.k[j][n_plane + i], ei); }
    er *= c.h; ei *= c.h;
    const double y2 = c.y[i] * c.y[i] + c.y[n_plane + i] * c.y[n_plane + i];
    const double n2 = c.yn[i] * c.yn[i] + c.yn[n_plane + i] * c.yn[n_plane + i];
    const double s = c.atol + c.rtol * sqrt(fmax(y2, n2));
    es += (er * er + ei * ei) / (s * s);
  }
  es = block_sum(es, red);
  if (threadIdx.x == 0) atomicAdd(c.out, es);
}
// out[e] = Tr(E_e rho) for a Hermitian rho held planar [2][rows][NP]: sum_ij E_ij conj(rho_ij)
__global__ void large_expect_kernel(const cplx* __restrict__ E, const double* __restrict__ Y, long long y_plane, int N, int NP,
                                    cplx* __restrict__ out) {
  __shared__ double red[34];
  const int e = blockIdx.y;
  const cplx* Em = E + (long long)e * N * N;
  cplx acc = cmake(0, 0);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / N;
    const int j = (int)(idx - i * N);
    const cplx ev = Em[idx];
    const double yr = Y[i * NP + j], yi = Y[y_plane + i * NP + j];
    acc.x += ev.x * yr + ev.y * yi;
    acc.y += ev.y * yr - ev.x * yi;
  }
  acc = block_sum_c(acc, red);
  if (threadIdx.x == 0) { atomicAdd(&out[e].x, acc.x); atomicAdd(&out[e].y, acc.y); }
}

// ---- chunk schedule of an operator (or of the union pattern of `nsrc` consecutive operators):
// block t lists the K chunks whose tile_rows x 16 window (rows row_begin + t*tile_rows ...) holds a non-zero
__global__ void large_sched_kernel(const double* __restrict__ base, long long op_stride, int nsrc, int NP, long long plane,
                                   long long row_begin, int rows_total, int tile_rows, int kchunks,
                                   unsigned short* __restrict__ idx, int* __restrict__ cnt) {
  extern __shared__ unsigned char sched_flags[];
  const int t = blockIdx.x;
  const long long r0 = row_begin + (long long)t * tile_rows;
  long long r1 = r0 + tile_rows;
  if (r1 > row_begin + rows_total) r1 = row_begin + rows_total;
  for (int kc = threadIdx.x; kc < kchunks; kc += blockDim.x) {
    bool nz = false;
    for (int sidx = 0; sidx < nsrc && !nz; sidx++) {
      const double* re = base + sidx * op_stride;
      const double* im = re + plane;
      for (long long r = r0; r < r1 && !nz; r++) {
        const double2* pr = reinterpret_cast<const double2*>(re + r * NP + 16 * kc);
        const double2* pi = reinterpret_cast<const double2*>(im + r * NP + 16 * kc);
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const double2 a = pr[c], b = pi[c];
          nz = nz || a.x != 0.0 || a.y != 0.0 || b.x != 0.0 || b.y != 0.0;
        }
      }
    }
    sched_flags[kc] = nz ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int kc = 0; kc < kchunks; kc++)
      if (sched_flags[kc]) idx[(long long)t * kchunks + n++] = (unsigned short)kc;
    cnt[t] = n;
  }
}
__global__ void large_pref_kernel(const int* __restrict__ cnt, int ntiles, int* __restrict__ pref) {
  if (blockIdx.x || threadIdx.x) return;
  int acc = 0;
  for (int t = 0; t < ntiles; t++) { pref[t] = acc; acc += cnt[t]; }
  pref[ntiles] = acc;
}

static inline int large_np(int N) { return (N + 15) / 16 * 16; }

template <bool BH>
static int launch_zgemm(const ZgemmArgs& a, int sms, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(zgemm_dmma_kernel<BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LargeGemmCfg::SMEM) != cudaSuccess)
      return SEQ_ERR_CUDA;
    configured = true;
  }
  const int grid = sms;   // stream-K: one CTA per SM (the listed iteration count lives on the device)
  if (a.M < 1 || a.Ncols < 1) return SEQ_OK;
  zgemm_dmma_kernel<BH><<<grid, LargeGemmCfg::NTHREADS, LargeGemmCfg::SMEM, st>>>(a);
  return SEQ_OK;
}

}  // namespace seq
