// SEQ_METHOD_DMMA: batched Lindblad integrator on the FP64 tensor cores (DMMA.8x8x4), rho resident.
//
// One CTA integrates one trajectory over the whole time span (Dormand-Prince 5(4), FSAL,
// CTA-uniform adaptive step control, fused spline evaluation and expectation reduction).
//
// Data layout
//  * N is padded to NP = 8*NT.  Matrices in shared memory are PLANAR (re plane, im plane), row
//    major with leading dimension LD = NP + 4 doubles: LD = 4 (mod 8) makes the A-fragment loads
//    (8 rows x 4 cols) and the B-fragment loads (4 rows x 8 cols) of mma.m8n8k4 bank-conflict free.
//  * shared memory holds four such matrices: Y (the RK stage state), T (scratch: -iGY, then L_l Y),
//    X0 (G(t), later odd L_l) and X1 (even L_l).  L_l are streamed from L2 with cp.async, double
//    buffered behind the products of the previous operator.
//  * each warp owns a 1 x TW strip of 8x8 complex output tiles; its accumulators (K: the RHS being
//    assembled, W: the running product) stay in registers for the whole RHS evaluation.
//  * the RK stage derivatives k1..k7, y and y_new are THREAD-PRIVATE streams in global memory,
//    stored in accumulator-fragment order ([slot][thread] double2, fully coalesced, L2 resident):
//    a thread only ever re-reads what it wrote, so the stage combinations need no barrier.
//
//  * operator TILE SPARSITY: the reference's operators are Kronecker-embedded ladder / number
//    operators, so most 8x4 fragments of G and of every L_l are exactly zero.  dmma_masks_kernel
//    records, per operator and 8-row tile, which of the NP/4 k-steps hold a non-zero fragment; the
//    product loops skip the others (warp-uniform branches, exact: only 0*x terms are dropped).
//    Dense operators have all bits set and run the full loops.
//
// RHS (rho Hermitian):  A = -i G rho,  G = H(t) - (i/2) sum_l r_l L_l^dag L_l
//                       rho' = A + A^dag + sum_l r_l L_l rho L_l^dag
// = (1 + 2 n_l) complex NP^3 products, each 3 real DMMA chains (Karatsuba form, see warp_gemm).
#pragma once
#include "common.cuh"

namespace seq {

// Optional region timers (build with -DSEQ_PROFILE_REGIONS; read back with cudaMemcpyFromSymbol via
// tools/region_prof.py): thread 0 of CTA 0 accumulates clock64() deltas per region of the step loop.
#ifdef SEQ_PROFILE_REGIONS
__device__ unsigned long long seq_prof[16];
#define PROF_DECL unsigned long long _pt = clock64();
#define PROF(r) do { if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long _n = clock64(); seq_prof[r] += _n - _pt; _pt = _n; } } while (0)
#define PROF_ARG , unsigned long long& _pt
#define PROF_PASS , _pt
#else
#define PROF_DECL
#define PROF(r)
#define PROF_ARG
#define PROF_PASS
#endif

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double dneg(double x) {
  return __longlong_as_double(__double_as_longlong(x) ^ 0x8000000000000000ull);  // LOP3, keeps the FP64 pipe free
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NT> struct DmmaCfg {
  static constexpr int NP = 8 * NT;
  static constexpr int LD = NP + 4;
  static constexpr int TW = (NT % 3 == 0) ? 3 : ((NT % 2 == 0) ? 2 : (NT == 5 ? 5 : 1));
  static constexpr int WPR = NT / TW;            // warps per tile row
  static constexpr int NW = NT * WPR;            // warps per CTA
  static constexpr int NTHREADS = NW * 32;
  static constexpr int PLANE = NP * LD;          // doubles per plane
  static constexpr int MAT = 2 * PLANE;          // doubles per planar matrix
  static constexpr int SLOTS = TW * 2;           // double2 per thread per state vector
  static constexpr int MAXL = 24;                // collapse operators whose masks / fragment lists fit the smem tables
  static constexpr int MAXF = 2 * NT * NT;        // 8x4 fragments per operator
  // 4 planar matrices + reduction scratch + coefficient values of the 7 stages + masks + fragment lists
  static constexpr size_t SMEM_BASE = (size_t)(4 * MAT + 34 + 7 * 64) * sizeof(double) + (size_t)(1 + MAXL) * 8 * sizeof(unsigned)
                                      + (size_t)(((1 + MAXL) * (MAXF + 8) + 15) / 16 * 16);
  // the rest of the SM's shared memory keeps the non-zero fragments of G0 and every Gk resident
  // (512 B per fragment), so the per-RHS G(t) assembly never leaves the SM when they fit
  static constexpr size_t SMEM_LIMIT = 227 * 1024;
  // (only where the CTA already owns the SM: smaller NT keep several CTAs per SM instead)
  static constexpr int GRES_FRAGS = (NT == 6) ? (int)((SMEM_LIMIT - SMEM_BASE) / 512) : 0;
  static constexpr size_t SMEM = SMEM_BASE + (size_t)GRES_FRAGS * 512;
};

// ---- prep: pack interleaved complex [N,N] into planar padded [2][NP][LD] (zero padded)
__global__ void dmma_pack_planar_kernel(const cplx* __restrict__ src, double* __restrict__ dst, int N, int NP, int LD) {
  int m = blockIdx.y;
  const cplx* s = src + (long long)m * N * N;
  double* d = dst + (long long)m * 2 * NP * LD;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NP * LD; idx += gridDim.x * blockDim.x) {
    int i = idx / LD, j = idx - i * LD;
    cplx v = cmake(0, 0);
    if (i < N && j < N) v = s[i * N + j];
    d[idx] = v.x;
    d[NP * LD + idx] = v.y;
  }
}

// ---- prep: E^T in accumulator-fragment order: Ef[e][slot][tid] = (E[c0][r], E[c0+1][r]) for the
// element (r, c0..c0+1) that thread `tid` owns in slot (t, plane)
template <int NT>
__global__ void dmma_pack_efrag_kernel(const cplx* __restrict__ E, double2* __restrict__ Ef, int N) {
  typedef DmmaCfg<NT> C;
  int e = blockIdx.x;
  int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tg = lane & 3;
  int tr = w / C::WPR, tc0 = (w % C::WPR) * C::TW;
  for (int t = 0; t < C::TW; t++) {
    int r = 8 * tr + g, c0 = 8 * (tc0 + t) + 2 * tg;
    cplx a = cmake(0, 0), b = cmake(0, 0);
    if (r < N && c0 < N) a = E[(long long)e * N * N + c0 * N + r];
    if (r < N && c0 + 1 < N) b = E[(long long)e * N * N + (c0 + 1) * N + r];
    Ef[((long long)e * C::SLOTS + t * 2 + 0) * C::NTHREADS + tid] = make_double2(a.x, b.x);
    Ef[((long long)e * C::SLOTS + t * 2 + 1) * C::NTHREADS + tid] = make_double2(a.y, b.y);
  }
}

// ---- prep: tile-sparsity masks.  masks[(op * 8) + tr] bit ks = 1 iff rows 8tr..8tr+7, columns
// 4ks..4ks+3 of packed operator `op` hold a non-zero (either plane).  Operator 0 is the union
// pattern of G0 and every Gk (the pattern of G(t) at any t), operators 1.. are the L_l.
// `blk`/`LD`/`PLANE` describe one planar image; `nimg` images per operator (1: resident layout,
// 4: the 2x2 block images of the streaming kernel, masks then indexed [op][img][tr]).
__global__ void dmma_masks_kernel(const double* __restrict__ G0p, const double* __restrict__ Gkp, int n_g,
                                  const double* __restrict__ Lp, int n_l, int NT, int LD, int nimg, unsigned* __restrict__ masks) {
  const int op = blockIdx.x;                 // 0: G union, 1..n_l: L
  const int PLANE = 8 * NT * LD, IMG = 2 * PLANE, nks = 2 * NT;
  for (int w = threadIdx.x; w < nimg * NT; w += blockDim.x) {
    const int img = w / NT, tr = w - img * NT;
    unsigned m = 0;
    const int nsrc = (op == 0) ? 1 + n_g : 1;
    for (int sidx = 0; sidx < nsrc; sidx++) {
      const double* base = (op == 0) ? (sidx == 0 ? G0p : Gkp + (long long)(sidx - 1) * nimg * IMG) : Lp + (long long)(op - 1) * nimg * IMG;
      base += (long long)img * IMG;
      for (int ks = 0; ks < nks; ks++) {
        bool nz = false;
        for (int r = 0; r < 8; r++)
          for (int c = 0; c < 4; c++) {
            const int off = (8 * tr + r) * LD + 4 * ks + c;
            nz = nz || base[off] != 0.0 || base[PLANE + off] != 0.0;
          }
        if (nz) m |= 1u << ks;
      }
    }
    masks[(op * nimg + img) * 8 + tr] = m;
  }
}

struct DmmaExtra {
  const double* G0p;   // planar padded G0 (shared)
  const double* Gkp;   // [n_g] planar padded
  const double* Lp;    // [n_l] planar padded
  const double2* Ef;   // fragment-ordered E^T
  const unsigned* masks;  // tile-sparsity masks (dmma_masks_kernel)
};

template <int NT>
struct DmmaState {
  typedef DmmaCfg<NT> C;
  double* Y; double* T; double* X0; double* X1; double* red; double* cval;
  const unsigned* masks;   // smem copy: [(1 + n_l)][8]
  const unsigned char* flist;   // [(1 + n_l)][MAXF + 8]: byte 0..3 = count (little endian), then (tr << 4 | ks) per non-zero fragment
  const double* gres;           // resident fragments of G0, Gk[0..n_g): [(1 + n_g)][nfrag][64] (nullptr: they do not fit)
  int tid, w, lane, g, tg, tr, tc0;
};

// acc += A(rows of tile-row tr) * B(cols of this warp's strip), A and B planar in shared memory.
// Complex product with THREE real DMMA chains per tile-step (Gauss/Karatsuba) instead of four:
//   P1 = Ar Br, P2 = Ai Bi, P3 = (Ar+Ai)(Br+Bi)      ->  Cr = P1 - P2,  Ci = P3 - P1 - P2
//   B = conj(L)^T:            P3 = (Ar+Ai)(Lr-Li)      ->  Cr = P1 + P2,  Ci = P3 - P1 + P2
// The operand sums run on the FP64 FMA pipe, which is otherwise idle while the tensor pipe is busy.
// mask[t] bit ks: execute k-step ks for output tile t (0 = the operator fragment is exactly zero).
template <int NT, bool BCONJT>
__device__ __forceinline__ void warp_gemm(double (&acc)[DmmaCfg<NT>::TW][2][2], const double* __restrict__ A,
                                          const double* __restrict__ Bm, int tr, int tc0, int g, int tg,
                                          const unsigned (&mask)[DmmaCfg<NT>::TW]) {
  typedef DmmaCfg<NT> C;
  const double* Are = A + (8 * tr + g) * C::LD + tg;
  const double* Aim = Are + C::PLANE;
  double P[C::TW][3][2];
#pragma unroll
  for (int t = 0; t < C::TW; t++)
#pragma unroll
    for (int i = 0; i < 3; i++) { P[t][i][0] = 0.0; P[t][i][1] = 0.0; }
  // Loops over the SET BITS of the masks (dynamic trip counts: real warp-uniform branches - a static
  // unrolled loop gets if-converted into predicated-off DMMAs that still occupy the tensor pipe).
  constexpr unsigned FULL = (1u << (C::NP / 4)) - 1u;
  unsigned all = FULL;
#pragma unroll
  for (int t = 0; t < C::TW; t++) all &= mask[t];
  if (all == FULL) {
    // dense operator rows: the statically unrolled loop (the compiler software-pipelines the fragment loads)
#pragma unroll
    for (int k0 = 0; k0 < C::NP; k0 += 4) {
      const double ar = Are[k0], ai = Aim[k0];
      const double as = ar + ai;
#pragma unroll
      for (int t = 0; t < C::TW; t++) {
        const int n0 = 8 * (tc0 + t);
        double br, bi, bs;
        if (!BCONJT) {
          const double* bp = Bm + (k0 + tg) * C::LD + n0 + g;
          br = bp[0]; bi = bp[C::PLANE];
          bs = br + bi;
        } else {
          const double* lp = Bm + (n0 + g) * C::LD + k0 + tg;   // B(k,n) = conj(L[n][k])
          br = lp[0]; bi = lp[C::PLANE];
          bs = br - bi;
        }
        dmma884(P[t][0][0], P[t][0][1], ar, br);
        dmma884(P[t][1][0], P[t][1][1], ai, bi);
        dmma884(P[t][2][0], P[t][2][1], as, bs);
      }
    }
  } else if (!BCONJT) {
    // the operator is the A operand: one mask (row tile tr) for the whole strip
    unsigned m = mask[0];
    while (m) {
      const int k0 = 4 * (__ffs(m) - 1);
      m &= m - 1;
      const double ar = Are[k0], ai = Aim[k0];
      const double as = ar + ai;
#pragma unroll
      for (int t = 0; t < C::TW; t++) {
        const double* bp = Bm + (k0 + tg) * C::LD + 8 * (tc0 + t) + g;
        const double br = bp[0], bi = bp[C::PLANE];
        const double bs = br + bi;
        dmma884(P[t][0][0], P[t][0][1], ar, br);
        dmma884(P[t][1][0], P[t][1][1], ai, bi);
        dmma884(P[t][2][0], P[t][2][1], as, bs);
      }
    }
  } else {
    // the operator is the B operand (conj-transpose): one mask per output tile (row tile tc0+t of L)
#pragma unroll
    for (int t = 0; t < C::TW; t++) {
      unsigned m = mask[t];
      const double* lrow = Bm + (8 * (tc0 + t) + g) * C::LD + tg;   // B(k,n) = conj(L[n][k])
      while (m) {
        const int k0 = 4 * (__ffs(m) - 1);
        m &= m - 1;
        const double ar = Are[k0], ai = Aim[k0];
        const double as = ar + ai;
        const double br = lrow[k0], bi = lrow[k0 + C::PLANE];
        const double bs = br - bi;
        dmma884(P[t][0][0], P[t][0][1], ar, br);
        dmma884(P[t][1][0], P[t][1][1], ai, bi);
        dmma884(P[t][2][0], P[t][2][1], as, bs);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < C::TW; t++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      if (!BCONJT) {
        acc[t][0][c] += P[t][0][c] - P[t][1][c];
        acc[t][1][c] += P[t][2][c] - P[t][0][c] - P[t][1][c];
      } else {
        acc[t][0][c] += P[t][0][c] + P[t][1][c];
        acc[t][1][c] += P[t][2][c] - P[t][0][c] + P[t][1][c];
      }
    }
}

// offset (doubles) inside a planar image of 16-byte chunk `w` (0..31) of fragment byte `fb` = (tr << 4 | ks):
// 2 planes x 8 rows x 2 chunks
template <int NT>
__device__ __forceinline__ int frag_chunk_off(unsigned fb, int w) {
  typedef DmmaCfg<NT> C;
  const int tr = fb >> 4, ks = fb & 15;
  return (w >> 4) * C::PLANE + (8 * tr + ((w >> 1) & 7)) * C::LD + 4 * ks + 2 * (w & 1);
}
template <int NT>
__device__ __forceinline__ int frag_count(const unsigned char* fl) {
  return (int)fl[0] | ((int)fl[1] << 8);
}

// stream the NON-ZERO fragments of collapse operator `op` (1-based list index) into `dst`: the
// other fragments of the buffer keep stale data that the masked product loops never read
template <int NT>
__device__ __forceinline__ void prefetch_L(const DmmaState<NT>& s, double* dst, const double* __restrict__ src, int op) {
  typedef DmmaCfg<NT> C;
  const unsigned char* fl = s.flist + op * (C::MAXF + 8);
  const int n = frag_count<NT>(fl) * 32;
  for (int c = s.tid; c < n; c += C::NTHREADS) {
    const int off = frag_chunk_off<NT>(fl[8 + (c >> 5)], c & 31);
    cp_async16(dst + off, src + off);
  }
  cp_async_commit();
}

// K (registers) <- RHS(t, Y in smem).  On exit no thread is still reading Y/T/X buffers is NOT
// guaranteed: callers must barrier before overwriting Y.
template <int NT>
__device__ __forceinline__ void dmma_rhs(const SolveParams& p, const DmmaExtra& x, DmmaState<NT>& s, const double* __restrict__ cv,
                                         double (&K)[DmmaCfg<NT>::TW][2][2] PROF_ARG) {
  typedef DmmaCfg<NT> C;
  const int nL = p.n_c + p.n_ctd;
  // stream the first collapse operator behind the G products
  if (nL > 0) prefetch_L<NT>(s, s.X1, x.Lp, 1);
  __syncthreads();  // every warp finished the previous RHS (X0/T free), Y written, cv visible
  PROF(1);
  // G(t) = G0 + sum_m c_m Gk[m]  -> X0, only on the fragments of the union pattern.  Operators are
  // taken two at a time with every 16-byte L2 load of the pair in flight together.
  if (s.gres) {
    // resident fragments: shared memory -> shared memory
    const unsigned char* fl = s.flist;
    const int n = frag_count<NT>(fl) * 32;
    const double2* gr = reinterpret_cast<const double2*>(s.gres);
    for (int c = s.tid; c < n; c += C::NTHREADS) {
      double2 acc = gr[c];
      for (int m = 0; m < p.n_g; m++) {
        const double2 a = gr[(m + 1) * n + c];
        acc.x = fma(cv[m], a.x, acc.x); acc.y = fma(cv[m], a.y, acc.y);
      }
      *reinterpret_cast<double2*>(s.X0 + frag_chunk_off<NT>(fl[8 + (c >> 5)], c & 31)) = acc;
    }
  } else {
    constexpr int CH = (C::MAXF * 32 + C::NTHREADS - 1) / C::NTHREADS;     // chunks per thread (dense worst case)
    const unsigned char* fl = s.flist;
    const int n = frag_count<NT>(fl) * 32;
    int off[CH];
    double2 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
      const int c = s.tid + i * C::NTHREADS;
      off[i] = (c < n) ? frag_chunk_off<NT>(fl[8 + (c >> 5)], c & 31) : -1;
    }
#pragma unroll
    for (int i = 0; i < CH; i++)
      acc[i] = (off[i] >= 0) ? *reinterpret_cast<const double2*>(x.G0p + off[i]) : make_double2(0.0, 0.0);
    for (int m = 0; m < p.n_g; m += 2) {
      const bool two = m + 1 < p.n_g;
      const double* g0 = x.Gkp + (long long)m * C::MAT;
      const double* g1 = g0 + C::MAT;
      const double c0 = cv[m], c1 = two ? cv[m + 1] : 0.0;
      double2 a0[CH], a1[CH];
#pragma unroll
      for (int i = 0; i < CH; i++) {
        a0[i] = (off[i] >= 0) ? *reinterpret_cast<const double2*>(g0 + off[i]) : make_double2(0.0, 0.0);
        a1[i] = (two && off[i] >= 0) ? *reinterpret_cast<const double2*>(g1 + off[i]) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int i = 0; i < CH; i++) {
        acc[i].x = fma(c0, a0[i].x, acc[i].x); acc[i].y = fma(c0, a0[i].y, acc[i].y);
        acc[i].x = fma(c1, a1[i].x, acc[i].x); acc[i].y = fma(c1, a1[i].y, acc[i].y);
      }
    }
#pragma unroll
    for (int i = 0; i < CH; i++)
      if (off[i] >= 0) *reinterpret_cast<double2*>(s.X0 + off[i]) = acc[i];
  }
  __syncthreads();
  PROF(2);
  double W[C::TW][2][2];
#pragma unroll
  for (int t = 0; t < C::TW; t++) { W[t][0][0] = W[t][0][1] = W[t][1][0] = W[t][1][1] = 0.0; }
  {
    unsigned mg[C::TW];
#pragma unroll
    for (int t = 0; t < C::TW; t++) mg[t] = s.masks[s.tr];
    warp_gemm<NT, false>(W, s.X0, s.Y, s.tr, s.tc0, s.g, s.tg, mg);
  }
  // A = -i W  -> T (planar): A_re = W_im, A_im = -W_re ; keep own A in K
#pragma unroll
  for (int t = 0; t < C::TW; t++) {
    int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + t) + 2 * s.tg;
    double are0 = W[t][1][0], are1 = W[t][1][1], aim0 = -W[t][0][0], aim1 = -W[t][0][1];
    *reinterpret_cast<double2*>(s.T + r * C::LD + c0) = make_double2(are0, are1);
    *reinterpret_cast<double2*>(s.T + C::PLANE + r * C::LD + c0) = make_double2(aim0, aim1);
    K[t][0][0] = are0; K[t][0][1] = are1; K[t][1][0] = aim0; K[t][1][1] = aim1;
  }
  __syncthreads();  // A complete; X0 (G) no longer needed
  PROF(3);
  // K += A^dag : K_re(i,j) += A_re(j,i) ; K_im(i,j) -= A_im(j,i)
#pragma unroll
  for (int t = 0; t < C::TW; t++) {
    int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + t) + 2 * s.tg;
    K[t][0][0] += s.T[(c0)*C::LD + r];
    K[t][0][1] += s.T[(c0 + 1) * C::LD + r];
    K[t][1][0] -= s.T[C::PLANE + (c0)*C::LD + r];
    K[t][1][1] -= s.T[C::PLANE + (c0 + 1) * C::LD + r];
  }
  for (int l = 0; l < nL; l++) {
    double* Lb = (l & 1) ? s.X0 : s.X1;
    cp_async_wait<0>();
    if (l == 0) PROF(4);
    __syncthreads();  // L_l landed for everyone; all reads of T and of the other X buffer are finished
    PROF(5);
    if (l + 1 < nL) prefetch_L<NT>(s, (l & 1) ? s.X1 : s.X0, x.Lp + (long long)(l + 1) * C::MAT, l + 2);
#pragma unroll
    for (int t = 0; t < C::TW; t++) { W[t][0][0] = W[t][0][1] = W[t][1][0] = W[t][1][1] = 0.0; }
    unsigned ma[C::TW], mb[C::TW];
#pragma unroll
    for (int t = 0; t < C::TW; t++) { ma[t] = s.masks[(l + 1) * 8 + s.tr]; mb[t] = s.masks[(l + 1) * 8 + s.tc0 + t]; }
    warp_gemm<NT, false>(W, Lb, s.Y, s.tr, s.tc0, s.g, s.tg, ma);
    const double r_l = (l < p.n_c) ? 1.0 : cv[p.n_ch + (l - p.n_c)];
#pragma unroll
    for (int t = 0; t < C::TW; t++) {
      int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + t) + 2 * s.tg;
      *reinterpret_cast<double2*>(s.T + r * C::LD + c0) = make_double2(r_l * W[t][0][0], r_l * W[t][0][1]);
      *reinterpret_cast<double2*>(s.T + C::PLANE + r * C::LD + c0) = make_double2(r_l * W[t][1][0], r_l * W[t][1][1]);
    }
    __syncthreads();  // T = r_l L_l Y complete
    PROF(6);
    warp_gemm<NT, true>(K, s.T, Lb, s.tr, s.tc0, s.g, s.tg, mb);
    PROF(7);
  }
}

template <int NT>
__global__ void __launch_bounds__(DmmaCfg<NT>::NTHREADS, 1)
solve_dmma_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ DmmaExtra x) {
  typedef DmmaCfg<NT> C;
  extern __shared__ __align__(16) double smem[];
  DmmaState<NT> s;
  s.Y = smem;
  s.T = smem + C::MAT;
  s.X0 = smem + 2 * C::MAT;
  s.X1 = smem + 3 * C::MAT;
  s.red = smem + 4 * C::MAT;
  s.cval = s.red + 34;
  {
    const int nops = 1 + p.n_c + p.n_ctd;
    unsigned* mk = reinterpret_cast<unsigned*>(s.cval + 7 * 64);
    unsigned char* fl = reinterpret_cast<unsigned char*>(mk + (1 + C::MAXL) * 8);
    for (int i = threadIdx.x; i < nops * 8; i += C::NTHREADS) mk[i] = x.masks[i];
    __syncthreads();
    // one thread per operator lists its non-zero fragments (tr << 4 | ks)
    for (int op = threadIdx.x; op < nops; op += C::NTHREADS) {
      unsigned char* l = fl + op * (C::MAXF + 8);
      int n = 0;
      for (int tr = 0; tr < NT; tr++)
        for (int ks = 0; ks < 2 * NT; ks++)
          if (mk[op * 8 + tr] & (1u << ks)) l[8 + n++] = (unsigned char)((tr << 4) | ks);
      l[0] = (unsigned char)(n & 255); l[1] = (unsigned char)(n >> 8);
    }
    s.masks = mk;
    s.flist = fl;
    __syncthreads();
    double* gres = smem + C::SMEM_BASE / sizeof(double);
    const int nfg = (int)fl[0] | ((int)fl[1] << 8);
    s.gres = nullptr;
    if ((1 + p.n_g) * nfg <= C::GRES_FRAGS) {
      for (int m = 0; m <= p.n_g; m++) {
        const double* src = (m == 0) ? x.G0p : x.Gkp + (long long)(m - 1) * C::MAT;
        for (int c = threadIdx.x; c < nfg * 32; c += C::NTHREADS) {
          const int off = frag_chunk_off<NT>(fl[8 + (c >> 5)], c & 31);
          *reinterpret_cast<double2*>(gres + ((long long)m * nfg * 32 + c) * 2) = *reinterpret_cast<const double2*>(src + off);
        }
      }
      s.gres = gres;
    }
  }
  s.tid = threadIdx.x; s.w = s.tid >> 5; s.lane = s.tid & 31; s.g = s.lane >> 2; s.tg = s.lane & 3;
  s.tr = s.w / C::WPR; s.tc0 = (s.w % C::WPR) * C::TW;
  const int b = blockIdx.x, N = p.N, tid = s.tid;
  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;

  // thread-private streams: y, yn, k1..k7 ; each SLOTS double2 per thread
  double2* ws = reinterpret_cast<double2*>(p.ws + (long long)b * p.ws_stride) + tid;
  constexpr int STREAM = C::SLOTS * C::NTHREADS;
  double2* sy = ws; double2* syn = ws + STREAM;
  double2* kp[7];
#pragma unroll
  for (int j = 0; j < 7; j++) kp[j] = ws + (2 + j) * STREAM;
#define SLOT(ptr, t, pl) (ptr)[((t)*2 + (pl)) * C::NTHREADS]

  // zero the Y buffer once (padding must stay zero), then load rho0 into registers/streams
  for (int c = tid; c < C::MAT; c += C::NTHREADS) s.Y[c] = 0.0;
  double yv[C::TW][2][2];
  {
    const cplx* y0 = p.state0 + (long long)b * N * N;
#pragma unroll
    for (int t = 0; t < C::TW; t++) {
      int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + t) + 2 * s.tg;
      cplx a = cmake(0, 0), bb = cmake(0, 0);
      if (r < N && c0 < N) a = y0[r * N + c0];
      if (r < N && c0 + 1 < N) bb = y0[r * N + c0 + 1];
      yv[t][0][0] = a.x; yv[t][0][1] = bb.x; yv[t][1][0] = a.y; yv[t][1][1] = bb.y;
      SLOT(sy, t, 0) = make_double2(a.x, bb.x);
      SLOT(sy, t, 1) = make_double2(a.y, bb.y);
    }
  }
  __syncthreads();

  // write a state held in registers into the smem stage buffer Y
  auto put_Y = [&](const double (&v)[C::TW][2][2]) {
#pragma unroll
    for (int t = 0; t < C::TW; t++) {
      int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + t) + 2 * s.tg;
      *reinterpret_cast<double2*>(s.Y + r * C::LD + c0) = make_double2(v[t][0][0], v[t][0][1]);
      *reinterpret_cast<double2*>(s.Y + C::PLANE + r * C::LD + c0) = make_double2(v[t][1][0], v[t][1][1]);
    }
  };
  auto store_stream = [&](double2* dst, const double (&v)[C::TW][2][2]) {
#pragma unroll
    for (int t = 0; t < C::TW; t++) {
      SLOT(dst, t, 0) = make_double2(v[t][0][0], v[t][0][1]);
      SLOT(dst, t, 1) = make_double2(v[t][1][0], v[t][1][1]);
    }
  };
  // v = y + h * sum_j a_j k_j  (thread-private).  Stream-outer / slot-inner: the 2*TW loads of one
  // stream are independent and issued together.
  auto combine = [&](double (&v)[C::TW][2][2], double h, int n, const double* __restrict__ a, double2* const* ks) {
    double2 acc[C::TW][2];
#pragma unroll
    for (int t = 0; t < C::TW; t++) { acc[t][0] = make_double2(0.0, 0.0); acc[t][1] = make_double2(0.0, 0.0); }
    // three streams per round: their 6*TW 16-byte L2 loads are issued together
    for (int j0 = 0; j0 < n; j0 += 3) {
      double2 kv[3][C::TW][2];
      double aj[3];
#pragma unroll
      for (int u = 0; u < 3; u++) {
        const bool on = j0 + u < n;
        aj[u] = on ? a[j0 + u] : 0.0;
        const double2* kj = ks[on ? j0 + u : j0];
#pragma unroll
        for (int t = 0; t < C::TW; t++) { kv[u][t][0] = SLOT(kj, t, 0); kv[u][t][1] = SLOT(kj, t, 1); }
      }
#pragma unroll
      for (int u = 0; u < 3; u++)
#pragma unroll
        for (int t = 0; t < C::TW; t++)
#pragma unroll
          for (int pl = 0; pl < 2; pl++) {
            acc[t][pl].x = fma(aj[u], kv[u][t][pl].x, acc[t][pl].x);
            acc[t][pl].y = fma(aj[u], kv[u][t][pl].y, acc[t][pl].y);
          }
    }
#pragma unroll
    for (int t = 0; t < C::TW; t++)
#pragma unroll
      for (int pl = 0; pl < 2; pl++) {
        double2 y0 = SLOT(sy, t, pl);
        v[t][pl][0] = fma(h, acc[t][pl].x, y0.x);
        v[t][pl][1] = fma(h, acc[t][pl].y, y0.y);
      }
  };

  // CTA-wide sum with ONE barrier: warp partials into one of two alternating slots, every thread adds
  // the NW partials itself in the same order (so the result - and the step decision - is CTA-uniform)
  int red_parity = 0;
  auto block_sum_1b = [&](double v) {
    v = warp_sum(v);
    double* slot = s.red + 16 * red_parity;
    red_parity ^= 1;
    if (s.lane == 0) slot[s.w] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < C::NW; i++) tot += slot[i];
    return tot;
  };

  StepCtl ctl;
  ctl_init(ctl, p);
  double K[C::TW][2][2];
  PROF_DECL

  for (int o = 0; o < p.n_out; o++) {
    if (o > 0) {
      const double t_target = p.t0 + p.dt * (double)p.out_idx[o];
      int nst = 0;
      double htry;
      bool landing, capped;
      while (ctl.status == SEQ_STATUS_OK && ctl_next(ctl, p, t_target, htry, landing, capped)) {
        const int pair = ctl.pair, S = RK_STAGES[pair];
        // spline values of every time-dependent term at every stage time of this step (one L2 round
        // trip per step instead of one per RHS); made visible by the barrier inside the first dmma_rhs
        for (int i = tid; i < S * p.n_g; i += C::NTHREADS) {
          const int st = i / p.n_g, m = i - st * p.n_g;
          const double ts = ctl.t + RK_C[pair][st] * htry;
          double v;
          if (m < p.n_ch) {
            v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, ts);
            if (cscale) v *= cscale[m];
          } else {
            v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, ts);
          }
          s.cval[st * 64 + m] = v;
        }
        PROF(15);
        for (int st = ctl.have_k1 ? 1 : 0; st < S; st++) {
          double v[C::TW][2][2];
          PROF(10);
          combine(v, htry, st, RK_A[pair][st], kp);    // st == S-1: v = y_new
          PROF(0);
          __syncthreads();                             // every warp is out of the previous RHS
          PROF(8);
          put_Y(v);
          if (st == S - 1) store_stream(syn, v);
          dmma_rhs<NT>(p, x, s, s.cval + st * 64, K PROF_PASS);
          store_stream(kp[st], K);
          PROF(9);
          ctl.nrhs++;
        }
        ctl.have_k1 = true;
        // error estimate (thread-private streams), RMS over the N*N physical elements
        double es = 0.0;
        {
          double2 e[C::TW][2];
#pragma unroll
          for (int tt = 0; tt < C::TW; tt++) { e[tt][0] = make_double2(0.0, 0.0); e[tt][1] = make_double2(0.0, 0.0); }
          for (int j0 = 0; j0 < S; j0 += 3) {
            double2 kv[3][C::TW][2];
            double ej[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const bool on = j0 + u < S;
              ej[u] = on ? RK_E[pair][j0 + u] : 0.0;
              const double2* kj = kp[on ? j0 + u : j0];
#pragma unroll
              for (int tt = 0; tt < C::TW; tt++) { kv[u][tt][0] = SLOT(kj, tt, 0); kv[u][tt][1] = SLOT(kj, tt, 1); }
            }
#pragma unroll
            for (int u = 0; u < 3; u++)
#pragma unroll
              for (int tt = 0; tt < C::TW; tt++)
#pragma unroll
                for (int pl = 0; pl < 2; pl++) {
                  e[tt][pl].x = fma(ej[u], kv[u][tt][pl].x, e[tt][pl].x);
                  e[tt][pl].y = fma(ej[u], kv[u][tt][pl].y, e[tt][pl].y);
                }
          }
#pragma unroll
          for (int tt = 0; tt < C::TW; tt++) {
            double2 yr = SLOT(sy, tt, 0), yi = SLOT(sy, tt, 1), nr = SLOT(syn, tt, 0), ni = SLOT(syn, tt, 1);
            double sc0 = p.atol + p.rtol * sqrt(fmax(yr.x * yr.x + yi.x * yi.x, nr.x * nr.x + ni.x * ni.x));
            double sc1 = p.atol + p.rtol * sqrt(fmax(yr.y * yr.y + yi.y * yi.y, nr.y * nr.y + ni.y * ni.y));
            double e0 = htry * htry * (e[tt][0].x * e[tt][0].x + e[tt][1].x * e[tt][1].x);
            double e1 = htry * htry * (e[tt][0].y * e[tt][0].y + e[tt][1].y * e[tt][1].y);
            es += e0 / (sc0 * sc0) + e1 / (sc1 * sc1);
          }
        }
        PROF(11);
        es = block_sum_1b(es);
        PROF(12);
        double err = sqrt(es / (double)(N * N));
        if (ctl_update(ctl, p, err, htry, landing, capped, t_target, nst)) {
          double2* tmp = sy; sy = syn; syn = tmp;
          tmp = kp[0]; kp[0] = kp[S - 1]; kp[S - 1] = tmp;
        }
        PROF(13);
      }
      if (ctl.status != SEQ_STATUS_OK) break;
    }
    // ---- outputs at sample o (thread-private reads of the y stream)
#pragma unroll
    for (int tt = 0; tt < C::TW; tt++) {
      double2 a = SLOT(sy, tt, 0), bb = SLOT(sy, tt, 1);
      yv[tt][0][0] = a.x; yv[tt][0][1] = a.y; yv[tt][1][0] = bb.x; yv[tt][1][1] = bb.y;
    }
    // Tr(E rho) for up to EC operators per round: thread partials -> the (idle) T buffer, one barrier,
    // then warp v adds column v and writes the result.  Ef holds E[c][r] for the owned (r,c).
    {
      constexpr int EC = (C::MAT / (2 * C::NTHREADS) < 8) ? C::MAT / (2 * C::NTHREADS) : 8;
      for (int e0 = 0; e0 < p.n_e; e0 += EC) {
        const int ne = (p.n_e - e0 < EC) ? p.n_e - e0 : EC;
        if (e0 > 0) __syncthreads();
        for (int ei_ = 0; ei_ < ne; ei_++) {
          const int e = e0 + ei_;
          cplx acc = cmake(0, 0);
#pragma unroll
          for (int tt = 0; tt < C::TW; tt++) {
            double2 er = x.Ef[((long long)e * C::SLOTS + tt * 2 + 0) * C::NTHREADS + tid];
            double2 ei = x.Ef[((long long)e * C::SLOTS + tt * 2 + 1) * C::NTHREADS + tid];
            // (er + i ei) * (yr + i yi)
            acc.x += er.x * yv[tt][0][0] - ei.x * yv[tt][1][0] + er.y * yv[tt][0][1] - ei.y * yv[tt][1][1];
            acc.y += er.x * yv[tt][1][0] + ei.x * yv[tt][0][0] + er.y * yv[tt][1][1] + ei.y * yv[tt][0][1];
          }
          s.T[(2 * ei_ + 0) * C::NTHREADS + tid] = acc.x;
          s.T[(2 * ei_ + 1) * C::NTHREADS + tid] = acc.y;
        }
        __syncthreads();
        for (int v = s.w; v < 2 * ne; v += C::NW) {
          double part = 0.0;
#pragma unroll
          for (int j = 0; j < C::NW; j++) part += s.T[v * C::NTHREADS + j * 32 + s.lane];
          part = warp_sum(part);
          if (s.lane == 0) {
            const int e = e0 + (v >> 1);
            long long off = ((long long)b * p.n_e + e) * p.n_out + o;
            if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((double*)p.expect)[2 * off + (v & 1)] = part;
            else if (!(v & 1)) ((double*)p.expect)[off] = part;
          }
        }
      }
      if (p.n_e > 0) __syncthreads();   // T is scratch again
    }
    PROF(14);
    if (p.flags & SEQ_FLAG_STORE_STATES) {
      cplx* dst = p.states + ((long long)b * p.n_out + o) * N * N;
#pragma unroll
      for (int tt = 0; tt < C::TW; tt++) {
        int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + tt) + 2 * s.tg;
        if (r < N && c0 < N) dst[r * N + c0] = cmake(yv[tt][0][0], yv[tt][1][0]);
        if (r < N && c0 + 1 < N) dst[r * N + c0 + 1] = cmake(yv[tt][0][1], yv[tt][1][1]);
      }
    }
  }
  {
    cplx* fin = p.final_state + (long long)b * N * N;
#pragma unroll
    for (int tt = 0; tt < C::TW; tt++) {
      double2 a = SLOT(sy, tt, 0), bb = SLOT(sy, tt, 1);
      int r = 8 * s.tr + s.g, c0 = 8 * (s.tc0 + tt) + 2 * s.tg;
      if (r < N && c0 < N) fin[r * N + c0] = cmake(a.x, bb.x);
      if (r < N && c0 + 1 < N) fin[r * N + c0 + 1] = cmake(a.y, bb.y);
    }
  }
  if (tid == 0) {
    p.status[b] = ctl.status;
    p.stats[b * 4 + 0] = ctl.nacc;
    p.stats[b * 4 + 1] = ctl.nrej;
    p.stats[b * 4 + 2] = ctl.nrhs;
    p.stats[b * 4 + 3] = 0;
  }
#undef SLOT
}

// ---- host side ---------------------------------------------------------------------------
static inline int dmma_nt(int N) { return (N + 7) / 8; }
static inline bool dmma_supported(int N, int max_smem) {
  int nt = dmma_nt(N);
  if (nt < 1 || nt > 6) return false;
  return DmmaCfg<6>::SMEM <= (size_t)max_smem;
}
// complex elements of per-trajectory scratch: 9 streams of SLOTS*NTHREADS double2
static inline size_t dmma_ws_elems(int N) {
  switch (dmma_nt(N)) {
    case 1: return 9 * (size_t)DmmaCfg<1>::SLOTS * DmmaCfg<1>::NTHREADS;
    case 2: return 9 * (size_t)DmmaCfg<2>::SLOTS * DmmaCfg<2>::NTHREADS;
    case 3: return 9 * (size_t)DmmaCfg<3>::SLOTS * DmmaCfg<3>::NTHREADS;
    case 4: return 9 * (size_t)DmmaCfg<4>::SLOTS * DmmaCfg<4>::NTHREADS;
    case 5: return 9 * (size_t)DmmaCfg<5>::SLOTS * DmmaCfg<5>::NTHREADS;
    default: return 9 * (size_t)DmmaCfg<6>::SLOTS * DmmaCfg<6>::NTHREADS;
  }
}
// doubles of packed-operator scratch the DMMA path needs
static inline size_t dmma_pack_doubles(int N, int n_g, int n_l, int n_e) {
  int nt = dmma_nt(N), NP = 8 * nt, LD = NP + 4;
  size_t mat = 2 * (size_t)NP * LD;
  size_t slots_threads = dmma_ws_elems(N) / 9;  // SLOTS*NTHREADS
  return (1 + (size_t)n_g + n_l) * mat + (size_t)n_e * slots_threads * 2 + 64 + (size_t)(1 + n_l) * 8;
}

template <int NT>
static int launch_dmma_nt(const SolveParams& p, const cplx* G0, const cplx* Gk, double* pack, cudaStream_t st, int* launches) {
  typedef DmmaCfg<NT> C;
  const int N = p.N, n_l = p.n_c + p.n_ctd;
  DmmaExtra x;
  double* G0p = pack;
  double* Gkp = G0p + C::MAT;
  double* Lp = Gkp + (size_t)p.n_g * C::MAT;
  double2* Ef = reinterpret_cast<double2*>(Lp + (size_t)n_l * C::MAT);
  dim3 blk(256);
  int gx = (C::NP * C::LD + 255) / 256;
  dmma_pack_planar_kernel<<<dim3(gx, 1), blk, 0, st>>>(G0, G0p, N, C::NP, C::LD);
  (*launches)++;
  if (p.n_g > 0) { dmma_pack_planar_kernel<<<dim3(gx, p.n_g), blk, 0, st>>>(Gk, Gkp, N, C::NP, C::LD); (*launches)++; }
  if (n_l > 0) { dmma_pack_planar_kernel<<<dim3(gx, n_l), blk, 0, st>>>(p.L, Lp, N, C::NP, C::LD); (*launches)++; }
  if (p.n_e > 0) { dmma_pack_efrag_kernel<NT><<<p.n_e, C::NTHREADS, 0, st>>>(p.E, Ef, N); (*launches)++; }
  unsigned* masks = reinterpret_cast<unsigned*>(Ef + (size_t)p.n_e * C::SLOTS * C::NTHREADS);
  dmma_masks_kernel<<<1 + n_l, 32, 0, st>>>(G0p, Gkp, p.n_g, Lp, n_l, NT, C::LD, 1, masks);
  (*launches)++;
  if (n_l > C::MAXL) return SEQ_ERR_UNSUPPORTED;
  x.G0p = G0p; x.Gkp = Gkp; x.Lp = Lp; x.Ef = Ef; x.masks = masks;
  cudaError_t rc = cudaFuncSetAttribute(solve_dmma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
  if (rc != cudaSuccess) return SEQ_ERR_CUDA;
  solve_dmma_kernel<NT><<<p.B, C::NTHREADS, C::SMEM, st>>>(p, x);
  (*launches)++;
  return SEQ_OK;
}

static inline int launch_dmma(const SolveParams& p, const cplx* G0, const cplx* Gk, double* pack, cudaStream_t st, int* launches) {
  switch (dmma_nt(p.N)) {
    case 1: return launch_dmma_nt<1>(p, G0, Gk, pack, st, launches);
    case 2: return launch_dmma_nt<2>(p, G0, Gk, pack, st, launches);
    case 3: return launch_dmma_nt<3>(p, G0, Gk, pack, st, launches);
    case 4: return launch_dmma_nt<4>(p, G0, Gk, pack, st, launches);
    case 5: return launch_dmma_nt<5>(p, G0, Gk, pack, st, launches);
    case 6: return launch_dmma_nt<6>(p, G0, Gk, pack, st, launches);
    default: return SEQ_ERR_UNSUPPORTED;
  }
}

}  // namespace seq
