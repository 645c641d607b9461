// SEQ_METHOD_DMMA_STREAM: FP64 tensor-core Lindblad integrator for 48 < N <= 96, where rho, G and
// the collapse operators no longer fit in shared memory together.
//
// The padded matrix (NPF = 2*BS, BS = 8*BT <= 48) is cut into 2x2 blocks.  Every matrix lives in L2
// as four "block images" (planar re/im, padded leading dimension BS+4 - the exact shared-memory
// layout of kernel_dmma.cuh), so a block moves L2 -> shared memory as a flat cp.async stream.  An RHS
// evaluation is a list of block products  C[I][J] += A[I][Kb] * B[Kb][J]  (8 for -iG rho, 16 per
// collapse operator) executed through a two-stage cp.async pipeline: while the 12 warps run the
// DMMA chains of task i out of stage i&1, the two operand blocks of task i+1 land in the other
// stage.  One __syncthreads per task.  Accumulators of a quadrant stay in registers over its two
// K-blocks; finished quadrants go to L2: T (scratch, block images) or the thread-private k-stream.
//
// Operator tile sparsity (see kernel_dmma.cuh): block products whose operator block is entirely zero
// are dropped from the task list (built once per launch from dmma_masks_kernel's masks), and inside
// the remaining products the zero 8x4 fragments are skipped.
//
// One CTA still integrates one trajectory over the whole time span; step control, spline
// evaluation, error norm, expectation values and state stores are fused exactly as in the
// resident kernel.
#pragma once
#include "kernel_dmma.cuh"

namespace seq {

template <int BT> struct StreamCfg {
  typedef DmmaCfg<BT> C;
  static constexpr int BS = C::NP;                 // block size (rows/cols)
  static constexpr int NPF = 2 * BS;               // padded full dimension
  static constexpr int BLK = C::MAT;               // doubles per block image
  static constexpr int FULL = 4 * BLK;             // doubles per full matrix
  static constexpr int QSLOTS = 4 * C::SLOTS;      // double2 per thread per state vector
  static constexpr int NTHREADS = C::NTHREADS;
  static constexpr int MAXL = 24;                  // collapse operators the task list has room for
  static constexpr int MAXTASK = 8 + 16 * MAXL;
  // + masks [(1 + MAXL)][4 images][8] (unsigned) + compact task list (unsigned short) + its length
  static constexpr size_t SMEM = (size_t)(4 * BLK + 34 + 7 * 64) * sizeof(double) + (size_t)(1 + MAXL) * 32 * sizeof(unsigned)
                                 + (size_t)MAXTASK * sizeof(unsigned short) + 16;
};
// task-list entry: id | flags
#define STK_FIRST 0x1000u   // zero the accumulators before this product
#define STK_LAST 0x2000u    // run the quadrant epilogue after it
#define STK_NOGEMM 0x4000u  // operator block is zero: epilogue only
#define STK_ADAG 0x8000u    // add A^dag (from T) into k before this task

// interleaved complex [N,N] -> four block images (zero padded)
__global__ void stream_pack_blocks_kernel(const cplx* __restrict__ src, double* __restrict__ dst, int N, int BS, int LD) {
  int m = blockIdx.y;
  const cplx* s = src + (long long)m * N * N;
  const int plane = BS * LD, blk = 2 * plane;
  double* d = dst + (long long)m * 4 * blk;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 4 * plane; idx += gridDim.x * blockDim.x) {
    int q = idx / plane, rem = idx - q * plane;
    int i = rem / LD, j = rem - i * LD;
    int gi = (q >> 1) * BS + i, gj = (q & 1) * BS + j;
    cplx v = cmake(0, 0);
    if (j < BS && gi < N && gj < N) v = s[gi * N + gj];
    d[q * blk + rem] = v.x;
    d[q * blk + plane + rem] = v.y;
  }
}

template <int BT>
__global__ void stream_pack_efrag_kernel(const cplx* __restrict__ E, double2* __restrict__ Ef, int N) {
  typedef StreamCfg<BT> S; typedef DmmaCfg<BT> C;
  int e = blockIdx.x;
  int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  int tr = w / C::WPR, tc0 = (w % C::WPR) * C::TW;
  for (int q = 0; q < 4; q++)
    for (int t = 0; t < C::TW; t++) {
      int r = (q >> 1) * S::BS + 8 * tr + g, c0 = (q & 1) * S::BS + 8 * (tc0 + t) + 2 * tg;
      cplx a = cmake(0, 0), b = cmake(0, 0);
      if (r < N && c0 < N) a = E[(long long)e * N * N + c0 * N + r];
      if (r < N && c0 + 1 < N) b = E[(long long)e * N * N + (c0 + 1) * N + r];
      long long base = ((long long)e * S::QSLOTS + q * C::SLOTS + t * 2) * S::NTHREADS + tid;
      Ef[base] = make_double2(a.x, b.x);
      Ef[base + S::NTHREADS] = make_double2(a.y, b.y);
    }
}

template <int BT>
__device__ __forceinline__ void stream_issue(double* stage, const double* __restrict__ Ablk, const double* __restrict__ Bblk, int tid) {
  typedef StreamCfg<BT> S;
  for (int c = tid; c < S::BLK / 2; c += S::NTHREADS) {
    cp_async16(stage + 2 * c, Ablk + 2 * c);
    cp_async16(stage + S::BLK + 2 * c, Bblk + 2 * c);
  }
  cp_async_commit();
}

template <int BT>
__global__ void __launch_bounds__(StreamCfg<BT>::NTHREADS, 1)
solve_dmma_stream_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ DmmaExtra x) {
  typedef StreamCfg<BT> S; typedef DmmaCfg<BT> C;
  extern __shared__ __align__(16) double smem[];
  double* red = smem + 4 * S::BLK;
  double* cval = red + 34;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  const int tr = w / C::WPR, tc0 = (w % C::WPR) * C::TW;
  const int b = blockIdx.x, N = p.N;
  const int nL = p.n_c + p.n_ctd;
  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;

  unsigned* masks = reinterpret_cast<unsigned*>(cval + 7 * 64);
  unsigned short* tlist = reinterpret_cast<unsigned short*>(masks + (1 + S::MAXL) * 32);
  int* ntl_p = reinterpret_cast<int*>(tlist + S::MAXTASK);
  for (int i = tid; i < (1 + nL) * 32; i += S::NTHREADS) masks[i] = x.masks[i];
  __syncthreads();
  if (tid == 0) {
    // compact list of the block products that have a non-zero operator block
    auto img_zero = [&](int op, int img) {
      unsigned m = 0;
      for (int t = 0; t < BT; t++) m |= masks[(op * 4 + img) * 8 + t];
      return m == 0;
    };
    int n = 0;
    bool adag_pending = false;
    for (int grp = 0; grp < 4 + 8 * nL; grp++) {          // one group = the two K-blocks of one quadrant product
      int id0 = 2 * grp, kind, l = -1, q;
      if (id0 < 8) { kind = 0; q = id0 >> 1; }
      else { int j = id0 - 8; l = j >> 4; int r = j & 15; kind = (r < 8) ? 1 : 2; q = (r & 7) >> 1; }
      if (id0 == 8) adag_pending = true;
      bool z[2];
      for (int kb = 0; kb < 2; kb++) {
        int img = (kind == 2) ? ((q & 1) * 2 + kb) : ((q >> 1) * 2 + kb);
        z[kb] = img_zero(kind == 0 ? 0 : l + 1, img);
      }
      unsigned extra = adag_pending ? STK_ADAG : 0u;
      if (z[0] && z[1]) {
        if (kind == 2) continue;                            // nothing to add to k
        tlist[n++] = (unsigned short)(id0 | STK_FIRST | STK_LAST | STK_NOGEMM | extra);   // T quadrant = 0
      } else if (z[0] || z[1]) {
        tlist[n++] = (unsigned short)((id0 + (z[0] ? 1 : 0)) | STK_FIRST | STK_LAST | extra);
      } else {
        tlist[n++] = (unsigned short)(id0 | STK_FIRST | extra);
        tlist[n++] = (unsigned short)((id0 + 1) | STK_LAST);
      }
      adag_pending = false;
    }
    *ntl_p = n;
  }
  __syncthreads();
  const int ntl = *ntl_p;

  // per-trajectory scratch: three block-image matrices, then 9 thread-private streams
  double* wsd = reinterpret_cast<double*>(p.ws + (long long)b * p.ws_stride);
  double* Gb = wsd; double* Yb = wsd + S::FULL; double* Tb = wsd + 2 * S::FULL;
  double2* st0 = reinterpret_cast<double2*>(wsd + 3 * S::FULL) + tid;
  constexpr int STREAM = S::QSLOTS * S::NTHREADS;
  double2* sy = st0; double2* syn = st0 + STREAM;
  double2* kp[7];
#pragma unroll
  for (int j = 0; j < 7; j++) kp[j] = st0 + (2 + j) * STREAM;
#define QSLOT(ptr, q, t, pl) (ptr)[((q)*C::SLOTS + (t)*2 + (pl)) * S::NTHREADS]
  // position of the (r, c0) pair this thread owns in quadrant q, tile t, inside a block image
  auto blk_off = [&](int t) { return (8 * tr + g) * C::LD + 8 * (tc0 + t) + 2 * tg; };

  // zero Y/T images once so the padding stays zero, load rho0 into the y stream
  for (int c = tid; c < S::FULL; c += S::NTHREADS) { Yb[c] = 0.0; Tb[c] = 0.0; }
  {
    const cplx* y0 = p.state0 + (long long)b * N * N;
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int t = 0; t < C::TW; t++) {
        int r = (q >> 1) * S::BS + 8 * tr + g, c0 = (q & 1) * S::BS + 8 * (tc0 + t) + 2 * tg;
        cplx a = cmake(0, 0), bb = cmake(0, 0);
        if (r < N && c0 < N) a = y0[r * N + c0];
        if (r < N && c0 + 1 < N) bb = y0[r * N + c0 + 1];
        QSLOT(sy, q, t, 0) = make_double2(a.x, bb.x);
        QSLOT(sy, q, t, 1) = make_double2(a.y, bb.y);
      }
  }
  __syncthreads();

  // Yb <- y + h sum_j a_j k_j (thread-private streams -> own positions of the block images);
  // optionally also keep the result in the stream `keep` (y_new)
  auto stage_state = [&](double h, int n, const double* __restrict__ a, double2* const* ks, double2* keep) {
    for (int q = 0; q < 4; q++) {
      double2 acc[C::TW][2];
#pragma unroll
      for (int t = 0; t < C::TW; t++) { acc[t][0] = make_double2(0.0, 0.0); acc[t][1] = make_double2(0.0, 0.0); }
      for (int j = 0; j < n; j++) {
        const double aj = a[j];
        const double2* kj = ks[j];
        double2 kv[C::TW][2];
#pragma unroll
        for (int t = 0; t < C::TW; t++) { kv[t][0] = QSLOT(kj, q, t, 0); kv[t][1] = QSLOT(kj, q, t, 1); }
#pragma unroll
        for (int t = 0; t < C::TW; t++)
#pragma unroll
          for (int pl = 0; pl < 2; pl++) {
            acc[t][pl].x = fma(aj, kv[t][pl].x, acc[t][pl].x);
            acc[t][pl].y = fma(aj, kv[t][pl].y, acc[t][pl].y);
          }
      }
#pragma unroll
      for (int t = 0; t < C::TW; t++) {
        double2 v[2];
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
          double2 y0 = QSLOT(sy, q, t, pl);
          v[pl] = make_double2(fma(h, acc[t][pl].x, y0.x), fma(h, acc[t][pl].y, y0.y));
          if (keep) QSLOT(keep, q, t, pl) = v[pl];
        }
        double* dst = Yb + q * S::BLK + blk_off(t);
        *reinterpret_cast<double2*>(dst) = v[0];
        *reinterpret_cast<double2*>(dst + C::PLANE) = v[1];
      }
    }
  };

  // kout (thread-private stream) <- RHS(t, Yb)
  auto rhs = [&](const double* __restrict__ cv, double2* kout) {
    __syncthreads();   // cval visible; Yb complete; every warp is out of the previous RHS
    {
      // G(t) = G0 + sum_m c_m Gk[m] over the block images, 8 independent 16-byte loads in flight
      const double2* g0 = reinterpret_cast<const double2*>(x.G0p);
      double2* dst = reinterpret_cast<double2*>(Gb);
      constexpr int NCHUNK = S::FULL / 2, UN = 8;
      for (int base = tid; base < NCHUNK; base += UN * S::NTHREADS) {
        double2 acc[UN];
#pragma unroll
        for (int i = 0; i < UN; i++) {
          int c = base + i * S::NTHREADS;
          acc[i] = (c < NCHUNK) ? g0[c] : make_double2(0.0, 0.0);
        }
        for (int m = 0; m < p.n_g; m++) {
          const double2* gm = reinterpret_cast<const double2*>(x.Gkp + (long long)m * S::FULL);
          const double cvm = cv[m];
          double2 a[UN];
#pragma unroll
          for (int i = 0; i < UN; i++) {
            int c = base + i * S::NTHREADS;
            a[i] = (c < NCHUNK) ? gm[c] : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int i = 0; i < UN; i++) { acc[i].x = fma(cvm, a[i].x, acc[i].x); acc[i].y = fma(cvm, a[i].y, acc[i].y); }
        }
#pragma unroll
        for (int i = 0; i < UN; i++) {
          int c = base + i * S::NTHREADS;
          if (c < NCHUNK) dst[c] = acc[i];
        }
      }
    }
    __syncthreads();   // Gb complete
    // task id -> operand block images, quadrant, K-block, kind (0: G*Y, 1: L*Y, 2: T*L^dag)
    auto task = [&](int i, const double*& A, const double*& B, int& q, int& kb, int& kind, int& l) {
      if (i < 8) {
        q = i >> 1; kb = i & 1; kind = 0; l = -1;
        A = Gb + ((q >> 1) * 2 + kb) * S::BLK;
        B = Yb + (kb * 2 + (q & 1)) * S::BLK;
      } else {
        int j = i - 8;
        l = j >> 4;
        int r = j & 15;
        const double* Ll = x.Lp + (long long)l * S::FULL;
        if (r < 8) {
          q = r >> 1; kb = r & 1; kind = 1;
          A = Ll + ((q >> 1) * 2 + kb) * S::BLK;
          B = Yb + (kb * 2 + (q & 1)) * S::BLK;
        } else {
          r -= 8;
          q = r >> 1; kb = r & 1; kind = 2;
          A = Tb + ((q >> 1) * 2 + kb) * S::BLK;
          B = Ll + ((q & 1) * 2 + kb) * S::BLK;   // L[J][Kb], used as conj-transpose
        }
      }
    };
    const double* A; const double* B; int q, kb, kind, l;
    {
      const unsigned e0 = tlist[0];
      task(e0 & 0xfff, A, B, q, kb, kind, l);
      if (!(e0 & STK_NOGEMM)) stream_issue<BT>(smem, A, B, tid); else cp_async_commit();
    }
    double W[C::TW][2][2];
    for (int i = 0; i < ntl; i++) {
      const unsigned ent = tlist[i];
      cp_async_wait<0>();
      __syncthreads();   // operands of task i landed; every warp finished task i-1 (its stage is free)
      if (ent & STK_ADAG) {
        // K += A1^dag, A1 = -iG rho sits complete in Tb:  K_re(r,c) += A_re(c,r), K_im(r,c) -= A_im(c,r)
        for (int qq = 0; qq < 4; qq++)
#pragma unroll
          for (int t = 0; t < C::TW; t++) {
            int rl = 8 * tr + g, cl = 8 * (tc0 + t) + 2 * tg;
            const double* src = Tb + ((qq & 1) * 2 + (qq >> 1)) * S::BLK;   // block (J, I)
            double2 kr = QSLOT(kout, qq, t, 0), ki = QSLOT(kout, qq, t, 1);
            kr.x += src[cl * C::LD + rl];
            kr.y += src[(cl + 1) * C::LD + rl];
            ki.x -= src[C::PLANE + cl * C::LD + rl];
            ki.y -= src[C::PLANE + (cl + 1) * C::LD + rl];
            QSLOT(kout, qq, t, 0) = kr;
            QSLOT(kout, qq, t, 1) = ki;
          }
        __syncthreads();   // all A^dag reads done before any epilogue overwrites Tb
      }
      const double* cur = smem + (i & 1) * 2 * S::BLK;
      const int cq = q, ckb = kb, ckind = kind, cl_ = l;
      if (i + 1 < ntl) {
        const unsigned en = tlist[i + 1];
        task(en & 0xfff, A, B, q, kb, kind, l);
        if (!(en & STK_NOGEMM)) stream_issue<BT>(smem + ((i + 1) & 1) * 2 * S::BLK, A, B, tid); else cp_async_commit();
      }
      if (ent & STK_FIRST) {
#pragma unroll
        for (int t = 0; t < C::TW; t++) { W[t][0][0] = W[t][0][1] = W[t][1][0] = W[t][1][1] = 0.0; }
      }
      if (!(ent & STK_NOGEMM)) {
        unsigned mk[C::TW];
        if (ckind == 2) {
          const int img = (cq & 1) * 2 + ckb;
#pragma unroll
          for (int t = 0; t < C::TW; t++) mk[t] = masks[((cl_ + 1) * 4 + img) * 8 + tc0 + t];
          warp_gemm<BT, true>(W, cur, cur + S::BLK, tr, tc0, g, tg, mk);
        } else {
          const int img = (cq >> 1) * 2 + ckb;
          const unsigned m1 = masks[((ckind == 0 ? 0 : cl_ + 1) * 4 + img) * 8 + tr];
#pragma unroll
          for (int t = 0; t < C::TW; t++) mk[t] = m1;
          warp_gemm<BT, false>(W, cur, cur + S::BLK, tr, tc0, g, tg, mk);
        }
      }
      if (ent & STK_LAST) {
        if (ckind == 0) {
#pragma unroll
          for (int t = 0; t < C::TW; t++) {
            double2 are = make_double2(W[t][1][0], W[t][1][1]), aim = make_double2(-W[t][0][0], -W[t][0][1]);
            double* dst = Tb + cq * S::BLK + blk_off(t);
            *reinterpret_cast<double2*>(dst) = are;
            *reinterpret_cast<double2*>(dst + C::PLANE) = aim;
            QSLOT(kout, cq, t, 0) = are;
            QSLOT(kout, cq, t, 1) = aim;
          }
        } else if (ckind == 1) {
          const double r_l = (cl_ < p.n_c) ? 1.0 : cv[p.n_ch + (cl_ - p.n_c)];
#pragma unroll
          for (int t = 0; t < C::TW; t++) {
            double* dst = Tb + cq * S::BLK + blk_off(t);
            *reinterpret_cast<double2*>(dst) = make_double2(r_l * W[t][0][0], r_l * W[t][0][1]);
            *reinterpret_cast<double2*>(dst + C::PLANE) = make_double2(r_l * W[t][1][0], r_l * W[t][1][1]);
          }
        } else {
#pragma unroll
          for (int t = 0; t < C::TW; t++) {
            double2 kr = QSLOT(kout, cq, t, 0), ki = QSLOT(kout, cq, t, 1);
            kr.x += W[t][0][0]; kr.y += W[t][0][1]; ki.x += W[t][1][0]; ki.y += W[t][1][1];
            QSLOT(kout, cq, t, 0) = kr;
            QSLOT(kout, cq, t, 1) = ki;
          }
        }
      }
    }
    if (nL == 0) {
      __syncthreads();
      for (int qq = 0; qq < 4; qq++)
#pragma unroll
        for (int t = 0; t < C::TW; t++) {
          int rl = 8 * tr + g, cl = 8 * (tc0 + t) + 2 * tg;
          const double* src = Tb + ((qq & 1) * 2 + (qq >> 1)) * S::BLK;
          double2 kr = QSLOT(kout, qq, t, 0), ki = QSLOT(kout, qq, t, 1);
          kr.x += src[cl * C::LD + rl];
          kr.y += src[(cl + 1) * C::LD + rl];
          ki.x -= src[C::PLANE + cl * C::LD + rl];
          ki.y -= src[C::PLANE + (cl + 1) * C::LD + rl];
          QSLOT(kout, qq, t, 0) = kr;
          QSLOT(kout, qq, t, 1) = ki;
        }
    }
  };

  int red_parity = 0;
  auto block_sum_1b = [&](double v) {
    v = warp_sum(v);
    double* slot = red + 16 * red_parity;
    red_parity ^= 1;
    if (lane == 0) slot[w] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < C::NW; i++) tot += slot[i];
    return tot;
  };

  StepCtl ctl;
  ctl_init(ctl, p);

  for (int o = 0; o < p.n_out; o++) {
    if (o > 0) {
      const double t_target = p.t0 + p.dt * (double)p.out_idx[o];
      int nst = 0;
      double htry;
      bool landing, capped;
      while (ctl.status == SEQ_STATUS_OK && ctl_next(ctl, p, t_target, htry, landing, capped)) {
        const int pair = ctl.pair, S_ = RK_STAGES[pair];
        // spline values of every time-dependent term at every stage time of this step
        for (int i = tid; i < S_ * p.n_g; i += S::NTHREADS) {
          const int st = i / p.n_g, m = i - st * p.n_g;
          const double ts = ctl.t + RK_C[pair][st] * htry;
          double v;
          if (m < p.n_ch) {
            v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, ts);
            if (cscale) v *= cscale[m];
          } else {
            v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, ts);
          }
          cval[st * 64 + m] = v;
        }
        for (int st = ctl.have_k1 ? 1 : 0; st < S_; st++) {
          __syncthreads();
          stage_state(htry, st, RK_A[pair][st], kp, st == S_ - 1 ? syn : nullptr);
          rhs(cval + st * 64, kp[st]);
          ctl.nrhs++;
        }
        ctl.have_k1 = true;
        double es = 0.0;
        for (int q = 0; q < 4; q++) {
          double2 e[C::TW][2];
#pragma unroll
          for (int tt = 0; tt < C::TW; tt++) { e[tt][0] = make_double2(0.0, 0.0); e[tt][1] = make_double2(0.0, 0.0); }
          for (int j = 0; j < S_; j++) {
            const double ej = RK_E[pair][j];
            const double2* kj = kp[j];
            double2 kv[C::TW][2];
#pragma unroll
            for (int tt = 0; tt < C::TW; tt++) { kv[tt][0] = QSLOT(kj, q, tt, 0); kv[tt][1] = QSLOT(kj, q, tt, 1); }
#pragma unroll
            for (int tt = 0; tt < C::TW; tt++)
#pragma unroll
              for (int pl = 0; pl < 2; pl++) {
                e[tt][pl].x = fma(ej, kv[tt][pl].x, e[tt][pl].x);
                e[tt][pl].y = fma(ej, kv[tt][pl].y, e[tt][pl].y);
              }
          }
#pragma unroll
          for (int tt = 0; tt < C::TW; tt++) {
            double2 yr = QSLOT(sy, q, tt, 0), yi = QSLOT(sy, q, tt, 1), nr = QSLOT(syn, q, tt, 0), ni = QSLOT(syn, q, tt, 1);
            double sc0 = p.atol + p.rtol * sqrt(fmax(yr.x * yr.x + yi.x * yi.x, nr.x * nr.x + ni.x * ni.x));
            double sc1 = p.atol + p.rtol * sqrt(fmax(yr.y * yr.y + yi.y * yi.y, nr.y * nr.y + ni.y * ni.y));
            double e0 = htry * htry * (e[tt][0].x * e[tt][0].x + e[tt][1].x * e[tt][1].x);
            double e1 = htry * htry * (e[tt][0].y * e[tt][0].y + e[tt][1].y * e[tt][1].y);
            es += e0 / (sc0 * sc0) + e1 / (sc1 * sc1);
          }
        }
        es = block_sum_1b(es);
        double err = sqrt(es / (double)(N * N));
        if (ctl_update(ctl, p, err, htry, landing, capped, t_target, nst)) {
          double2* tmp = sy; sy = syn; syn = tmp;
          tmp = kp[0]; kp[0] = kp[S_ - 1]; kp[S_ - 1] = tmp;
        }
      }
      if (ctl.status != SEQ_STATUS_OK) break;
    }
    for (int e = 0; e < p.n_e; e++) {
      cplx acc = cmake(0, 0);
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int tt = 0; tt < C::TW; tt++) {
          long long base = ((long long)e * S::QSLOTS + q * C::SLOTS + tt * 2) * S::NTHREADS + tid;
          double2 er = x.Ef[base], ei = x.Ef[base + S::NTHREADS];
          double2 yr = QSLOT(sy, q, tt, 0), yi = QSLOT(sy, q, tt, 1);
          acc.x += er.x * yr.x - ei.x * yi.x + er.y * yr.y - ei.y * yi.y;
          acc.y += er.x * yi.x + ei.x * yr.x + er.y * yi.y + ei.y * yr.y;
        }
      acc.x = block_sum_1b(acc.x);
      acc.y = block_sum_1b(acc.y);
      if (tid == 0) {
        long long off = ((long long)b * p.n_e + e) * p.n_out + o;
        if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = acc;
        else ((double*)p.expect)[off] = acc.x;
      }
    }
    if (p.flags & SEQ_FLAG_STORE_STATES) {
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N * N;
      for (int q = 0; q < 4; q++)
#pragma unroll
        for (int tt = 0; tt < C::TW; tt++) {
          int r = (q >> 1) * S::BS + 8 * tr + g, c0 = (q & 1) * S::BS + 8 * (tc0 + tt) + 2 * tg;
          double2 yr = QSLOT(sy, q, tt, 0), yi = QSLOT(sy, q, tt, 1);
          if (r < N && c0 < N) dst[r * N + c0] = cmake(yr.x, yi.x);
          if (r < N && c0 + 1 < N) dst[r * N + c0 + 1] = cmake(yr.y, yi.y);
        }
    }
  }
  {
    cplx* fin = p.final_state + (long long)b * N * N;
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int tt = 0; tt < C::TW; tt++) {
        int r = (q >> 1) * S::BS + 8 * tr + g, c0 = (q & 1) * S::BS + 8 * (tc0 + tt) + 2 * tg;
        double2 yr = QSLOT(sy, q, tt, 0), yi = QSLOT(sy, q, tt, 1);
        if (r < N && c0 < N) fin[r * N + c0] = cmake(yr.x, yi.x);
        if (r < N && c0 + 1 < N) fin[r * N + c0 + 1] = cmake(yr.y, yi.y);
      }
  }
  if (tid == 0) {
    p.status[b] = ctl.status;
    p.stats[b * 4 + 0] = ctl.nacc;
    p.stats[b * 4 + 1] = ctl.nrej;
    p.stats[b * 4 + 2] = ctl.nrhs;
    p.stats[b * 4 + 3] = 0;
  }
#undef QSLOT
}

// ---- host side ---------------------------------------------------------------------------
static inline int stream_bt(int N) { int nt = (N + 7) / 8; return (nt + 1) / 2; }   // tiles per block
static inline bool stream_supported(int N, int max_smem) {
  int bt = stream_bt(N);
  return N > 48 && bt >= 4 && bt <= 6 && StreamCfg<6>::SMEM <= (size_t)max_smem;   // (n_l <= MAXL is checked at launch)
}
template <int BT> static size_t stream_ws_elems_bt() {
  typedef StreamCfg<BT> S;
  // 3 block-image matrices (doubles) + 9 streams of QSLOTS*NTHREADS double2, in complex (16 B) units
  return (3 * (size_t)S::FULL + 1) / 2 + 9 * (size_t)S::QSLOTS * S::NTHREADS + 8;
}
static inline size_t stream_ws_elems(int N) {
  switch (stream_bt(N)) {
    case 4: return stream_ws_elems_bt<4>();
    case 5: return stream_ws_elems_bt<5>();
    default: return stream_ws_elems_bt<6>();
  }
}
static inline size_t stream_pack_doubles(int N, int n_g, int n_l, int n_e) {
  int bt = stream_bt(N), BS = 8 * bt, LD = BS + 4;
  size_t full = 4 * 2 * (size_t)BS * LD;
  size_t qslots_threads = (stream_ws_elems(N) - (3 * full + 1) / 2 - 8) / 9;
  return (1 + (size_t)n_g + n_l) * full + (size_t)n_e * qslots_threads * 2 + 64 + (size_t)(1 + n_l) * 32;
}

template <int BT>
static int launch_stream_bt(const SolveParams& p, const cplx* G0, const cplx* Gk, double* pack, cudaStream_t st, int* launches) {
  typedef StreamCfg<BT> S; typedef DmmaCfg<BT> C;
  const int N = p.N, n_l = p.n_c + p.n_ctd;
  double* G0p = pack;
  double* Gkp = G0p + S::FULL;
  double* Lp = Gkp + (size_t)p.n_g * S::FULL;
  double2* Ef = reinterpret_cast<double2*>(Lp + (size_t)n_l * S::FULL);
  dim3 blk(256);
  int gx = (4 * C::PLANE + 255) / 256;
  stream_pack_blocks_kernel<<<dim3(gx, 1), blk, 0, st>>>(G0, G0p, N, S::BS, C::LD);
  (*launches)++;
  if (p.n_g > 0) { stream_pack_blocks_kernel<<<dim3(gx, p.n_g), blk, 0, st>>>(Gk, Gkp, N, S::BS, C::LD); (*launches)++; }
  if (n_l > 0) { stream_pack_blocks_kernel<<<dim3(gx, n_l), blk, 0, st>>>(p.L, Lp, N, S::BS, C::LD); (*launches)++; }
  if (p.n_e > 0) { stream_pack_efrag_kernel<BT><<<p.n_e, S::NTHREADS, 0, st>>>(p.E, Ef, N); (*launches)++; }
  if (n_l > S::MAXL) return SEQ_ERR_UNSUPPORTED;
  unsigned* masks = reinterpret_cast<unsigned*>(Ef + (size_t)p.n_e * S::QSLOTS * S::NTHREADS);
  dmma_masks_kernel<<<1 + n_l, 32, 0, st>>>(G0p, Gkp, p.n_g, Lp, n_l, BT, C::LD, 4, masks);
  (*launches)++;
  DmmaExtra x;
  x.G0p = G0p; x.Gkp = Gkp; x.Lp = Lp; x.Ef = Ef; x.masks = masks;
  cudaError_t rc = cudaFuncSetAttribute(solve_dmma_stream_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM);
  if (rc != cudaSuccess) return SEQ_ERR_CUDA;
  solve_dmma_stream_kernel<BT><<<p.B, S::NTHREADS, S::SMEM, st>>>(p, x);
  (*launches)++;
  return SEQ_OK;
}

static inline int launch_stream(const SolveParams& p, const cplx* G0, const cplx* Gk, double* pack, cudaStream_t st, int* launches) {
  switch (stream_bt(p.N)) {
    case 4: return launch_stream_bt<4>(p, G0, Gk, pack, st, launches);
    case 5: return launch_stream_bt<5>(p, G0, Gk, pack, st, launches);
    case 6: return launch_stream_bt<6>(p, G0, Gk, pack, st, launches);
    default: return SEQ_ERR_UNSUPPORTED;
  }
}

}  // namespace seq
