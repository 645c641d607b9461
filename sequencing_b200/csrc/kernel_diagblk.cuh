// SEQ_METHOD_DIAG, block-local variant ("diagblk"): the operators-by-diagonals formulation of kernel_diag.cuh with
// ONE tensor factor of the Hilbert space kept inside a thread.
//
// The reference's systems are tensor products of modes (system.py:399-445) and every drive / collapse operator acts
// on ONE mode (modes.py:125-207, 719-753): in the product basis i = (hi, a, lo) with a the index of a small "local"
// mode of dimension D at stride s, an operator of that mode is a diagonal at offset da * s that never changes (hi, lo).
// A thread therefore owns the whole D x D block { rho[(r, a), (r', b)] } of ONE pair of outer indices r = (hi, lo),
// r' = (hi', lo') - and every term of the local mode (the qubit drive -i[c_x(t) (a + a^dag) + c_y(t) i(a^dag - a), rho],
// its decay / excitation pairs) reads the thread's OWN registers: no shared-memory gather, no address, no zero guard.
// Only terms of the other modes gather from the stage matrix Y in shared memory (C2: 1 gather per element for the
// cavity decay pair instead of 6; the qubit drive's four gathers and the qubit decay's gather are register moves).
// Which diagonals are local is decided on the DATA, not on a hint: `diagblk_cross_kernel` flags every diagonal that
// has a non-zero leaving its D s block for a candidate (s, D); such diagonals simply stay gathered terms.
//
// Ownership: rho Hermitian -> outer pairs (r, r') of the cyclic band r' = r .. r + R/2 (mod R), R = N / D; one thread
// per pair, `lpr` (8 or 16 ..) lanes per outer row so that a quarter-warp's gathers / stores are 8 consecutive complex
// numbers.  The pair r' = r holds a full (redundantly Hermitian) D x D block.  Owners write their elements and the
// mirrored ones into Y.  C2 (transmon(3) x cavity(15)): R = 15, 120 threads x 9 elements - small CTAs, 2-3 resident
// per SM, each with its own barriers, so one trajectory's serial section (error sum -> controller) is hidden behind
// the other trajectories' arithmetic instead of behind a speculating controller warp.
// Same equations, same stepper, same controller (StepCtl) as every other kernel: identical step decisions.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "kernel_diag.cuh"

namespace seq {

#define BLK_MAXD 4
#define BLK_ND (2 * BLK_MAXD - 1)
#define BLK_MAXLL 3       // local collapse pairs per shift

struct BlkOps {
  int D, s, R, lpr;         // local dimension and stride, outer dimension, lanes per outer row
  int SL, SR;               // byte strides of the local row index (s ld 16) and column index (s 16) inside Y / a table
  int gl_tab[BLK_ND];       // local G diagonal with shift da (index da + BLK_MAXD - 1): table inside a Gt buffer, or -1
  int nLl[BLK_ND];          // local collapse pairs (da, da)
  int ll_a[BLK_ND][BLK_MAXLL], ll_b[BLK_ND][BLK_MAXLL], ll_r[BLK_ND][BLK_MAXLL];
  int kslots;               // stage-derivative slots in shared memory (the remaining ones are global-memory streams)
  int wslot;                // static weights: 0 registers, 1 shared memory (behind the slots), 2 one global table for the batch
  int mirror;               // owners also write the mirrored (conjugated) block into Y: needed unless every gather stays inside the owned band
};

// flags[(c * n_ops + op) * (2N - 1) + o + N - 1] = 1 if operator `op` (numbering of diag_flags_kernel) has a non-zero on
// diagonal o = da * s_c, |da| < D_c, that leaves its block of D_c * s_c indices (then the diagonal is not local);
// flags[((n_cands + c) * n_ops + op) * (2N - 1) + ..] = 1 if diagonal o has a non-zero whose row and column differ in the
// local index (a gathered term on such a diagonal does not move every outer pair by the same shift)
struct BlkCands { int n; int s[8], D[8]; };
__global__ void diagblk_cross_kernel(const __grid_constant__ BlkCands cands, const cplx* __restrict__ G0, const cplx* __restrict__ Gk, int n_g,
                                     const cplx* __restrict__ L, int N, int* __restrict__ flags) {
  const int op = blockIdx.y, n_ops = gridDim.y, c = blockIdx.z;
  const int s = cands.s[c], D = cands.D[c];
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
    if (!nz) continue;
    const int i = (int)(t / N), j = (int)(t - (long long)i * N), o = j - i;
    // second half of `flags`: the diagonal changes the LOCAL index somewhere (then its outer shift is not one constant)
    if ((i / s) % D != (j / s) % D) flags[((long long)(gridDim.z + c) * n_ops + op) * (2 * N - 1) + (o + N - 1)] = 1;
    if (o % s != 0 || abs(o / s) >= D) continue;
    if (i / (D * s) != j / (D * s)) flags[((long long)c * n_ops + op) * (2 * N - 1) + (o + N - 1)] = 1;
  }
}

// conj(z) with the sign flipped on the integer pipe (the FP64 pipe is the one this kernel is short of)
__device__ __forceinline__ cplx blk_conj(const cplx& z) {
  return cmake(z.x, __hiloint2double(__double2hiint(z.y) ^ (int)0x80000000, __double2loint(z.y)));
}

// compiler-level memory fence: loads are not hoisted across it (bounds the operands in flight, i.e. the registers)
__device__ __forceinline__ void blk_fence() { asm volatile("" ::: "memory"); }

// The stage loop is ROLLED (one copy of the right-hand side in the instruction stream: the statically unrolled loop of
// kernel_diag.cuh puts 12 copies of it there, 400 KB of SASS for 9 elements per thread, and four warps per CTA then
// wait for instruction fetches a third of the time - ncu no_instruction 4.1 of 11.6 cycles per issue).  A rolled loop
// cannot index registers by stage, so the stage derivatives live in thread-private slots [slot][e][tid] (conflict-free
// 16-byte accesses): the first bl.kslots in shared memory, the rest (k5 / k6, used by the 5(4) pair only) in global
// memory streams that stay in L2.  K1REG keeps k1 - the FSAL derivative every stage input needs - in registers.
template <int D, bool K1REG, bool WREG, bool GUARD, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
solve_diagblk_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ DiagOps sp, const __grid_constant__ BlkOps bl) {
  extern __shared__ __align__(16) unsigned char smem_diag[];
#ifdef SEQ_PROFILE_REGIONS
  const unsigned long long prof_c0 = clock64();
  unsigned long long prof_g0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_g0));
#endif
  constexpr int C = (int)sizeof(cplx);
  constexpr int EPT = D * D;
  const int N = p.N, ld = sp.ld, n_g = p.n_g;
  constexpr int nth = MAXT;      // launched with exactly MAXT threads: slot strides are immediates
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nw = (nth + 31) >> 5;
  const int b = blockIdx.x;
  const DiagLayout L = diag_layout(N, n_g, p.n_t, sp);
  double* cv = (double*)(smem_diag + L.off_cv);
  double* red = (double*)(smem_diag + L.off_red);
  DiagCtrl* ctrl = (DiagCtrl*)(smem_diag + L.off_ctrl);
  const int SL = bl.SL, SR = bl.SR;

  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;

  // ---- stage shared tables: zero guard bands / Y padding, collapse diagonals, G diagonals, E, spline tables
  for (int q = tid; q < L.off_Gt / 16; q += nth) ((cplx*)smem_diag)[q] = cmake(0, 0);
  for (int q = tid; q < sp.nT * N; q += nth) ((cplx*)(smem_diag + L.off_lam))[q] = sp.lam[q];
  const int gsz = sp.nD * N;
  if (sp.stage_g) {
    cplx* gs = (cplx*)(smem_diag + L.off_g);
    for (int q = tid; q < gsz; q += nth) gs[q] = sp.g0[q];
    for (int q = tid; q < n_g * gsz; q += nth) gs[gsz + q] = sp.gk[q];
  }
  const unsigned* eij = sp.e_ij;
  const cplx* eval_ = sp.e_val;
  if (sp.stage_e) {
    cplx* ev_s = (cplx*)(smem_diag + L.off_e);
    unsigned* ei_s = (unsigned*)(ev_s + (sp.nnzE > 0 ? sp.nnzE : 1));
    // staged lists hold RESOLVED locations: element index of rho_ji inside a Y buffer, bit 31 = "stored as its mirror
    // rho_ij, conjugate it" (without mirror writes only the owned orientation of an outer pair exists in Y)
    for (int q = tid; q < sp.nnzE; q += nth) {
      ev_s[q] = sp.e_val[q];
      const unsigned code = sp.e_ij[q];
      int row = (int)(code >> 16), col = (int)(code & 0xffffu);      // rho[row][col] = rho_ji
      unsigned flag = 0;
      if (!bl.mirror) {
        const int R = bl.R, s = bl.s, Ds = D * s;
        const int r1 = (row / Ds) * s + row % s, r2 = (col / Ds) * s + col % s;
        const int dd = (r2 - r1 + R) % R;
        const bool owned = dd < (R + 1) / 2 || ((R & 1) == 0 && dd == R / 2 && r1 < R / 2);
        if (!owned) { const int t = row; row = col; col = t; flag = 0x80000000u; }
      }
      ei_s[q] = (unsigned)(row * ld + col) | flag;
    }
    eij = ei_s; eval_ = ev_s;
  }
  if (sp.stage_tab) {
    double2* ts = (double2*)(smem_diag + L.off_tab);
    for (int q = tid; q < p.n_ch * p.n_t; q += nth) ts[q] = tabc[q];
    for (int q = tid; q < p.n_ctd * p.n_t; q += nth) ts[p.n_ch * p.n_t + q] = tabr[q];
    tabc = ts; tabr = ts + p.n_ch * p.n_t;
  }

  // ---- the owned outer pair (r, r2): byte offsets of Y[(r,0),(r2,0)] (ybase), of its mirror Y[(r2,0),(r,0)] (mbase) and
  // of entries (r,0) / (r2,0) of a diagonal table (oi / oj); local indices a, b add a SL + b SR (Y) or a SR, b SR (tables)
  int ybase, mbase, oi, oj;
  bool valid, isdiag;
  {
    const int R = bl.R, s = bl.s;
    const int r_raw = tid / bl.lpr, dpos = tid - r_raw * bl.lpr;
    valid = r_raw < R && dpos < R / 2 + 1 && !((R & 1) == 0 && dpos == R / 2 && r_raw >= R / 2);
    const int r = valid ? r_raw : 0;
    int r2 = r + (valid ? dpos : 0);
    if (r2 >= R) r2 -= R;
    isdiag = valid && dpos == 0;
    const int bi = (r / s) * D * s + r % s, bj = (r2 / s) * D * s + r2 % s;
    oi = bi * C; oj = bj * C;
    ybase = (bi * ld + bj) * C;
    mbase = (bj * ld + bi) * C;
    asm volatile("" : "+r"(ybase), "+r"(mbase), "+r"(oi), "+r"(oj));
  }
  cplx ys[EPT], kout[EPT], y[EPT];
  cplx W[WREG ? EPT : 1];
  cplx k1[K1REG ? EPT : 1];
  // slot of stage derivative j: K1REG ? j - 1 : j
  cplx* ksm = (cplx*)(smem_diag + L.off_w);
  cplx* kgl = p.ws + (long long)blockIdx.x * p.ws_stride;
  cplx* wst = bl.wslot == 1 ? ksm + (long long)bl.kslots * EPT * nth : p.ws + (long long)gridDim.x * p.ws_stride;   // static weights (global: one table for the batch)
  const int kslots = bl.kslots;
  {
    const cplx* y0 = p.state0 + (long long)b * N * N;
    const int bi = oi / C, bj = oj / C, s = bl.s;
#pragma unroll
    for (int a = 0; a < D; a++) {
      const cplx gi = sp.gdiag[bi + a * s];
#pragma unroll
      for (int c2 = 0; c2 < D; c2++) {
        const int e = a * D + c2;
        const int i = bi + a * s, j = bj + c2 * s;
        ys[e] = valid ? y0[i * N + j] : cmake(0, 0);
        y[e] = ys[e];
        kout[e] = cmake(0, 0);
        if (K1REG) k1[K1REG ? e : 0] = cmake(0, 0);
        // static own-element weight: -i (G0_ii - conj(G0_jj)) + sum over constant-rate pairs lam_a[i] conj(lam_b[j])
        const cplx gj = sp.gdiag[j];
        cplx w = cmake(gi.y + gj.y, -(gi.x - gj.x));
        for (int t = 0; t < sp.nLs; t++) {
          const cplx la = sp.lam[sp.ls_a[t] / C + i], lc = sp.lam[sp.ls_b[t] / C + j];
          w.x += la.x * lc.x + la.y * lc.y;
          w.y += la.y * lc.x - la.x * lc.y;
        }
        if (WREG) W[WREG ? e : 0] = w; else wst[e * nth + tid] = w;
      }
    }
  }
  __syncthreads();   // guard bands zeroed before anything is written into Y

  int yb = 0, gb = 0, tog = 0, ycur = 0;
  PROF_DECL
  // owner writes its block and (off the outer diagonal) the mirrored block
  const bool wmirror = !isdiag && bl.mirror != 0;
  auto write_Y = [&](int boff, const cplx (&v)[EPT]) {
    if (valid) {
      // (offsets laundered through an empty asm: derived addresses are recomputed here instead of living in registers
      // across the whole step loop as loop invariants)
      int o1 = boff + ybase, o2 = boff + mbase, sl = SL, sr = SR;
      asm volatile("" : "+r"(o1), "+r"(o2), "+r"(sl), "+r"(sr));
      unsigned char* yo = smem_diag + o1;
      unsigned char* ym = smem_diag + o2;
      const int SL = sl, SR = sr;
#pragma unroll
      for (int a = 0; a < D; a++)
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) *(cplx*)(yo + a * SL + c2 * SR) = v[a * D + c2];
      if (bl.mirror) {      // (a CTA-uniform branch: predicated-off stores would still take their issue slots)
#pragma unroll
        for (int a = 0; a < D; a++)
#pragma unroll
          for (int c2 = 0; c2 < D; c2++)
            if (wmirror) *(cplx*)(ym + c2 * SL + a * SR) = blk_conj(v[a * D + c2]);
      }
    }
  };
  // G(t) diagonals of one stage; two copies of the loop so that the staged tables are read with shared-memory loads
  // (one pointer that may be either makes every load a generic one)
  auto build_Gt = [&](cplx* Gb, const double* cvs, int st) {
    if (sp.stage_g) {
      const cplx* gs = (const cplx*)(smem_diag + L.off_g);
#pragma unroll 1
      for (int q = tid; q < gsz; q += nth) {
        cplx g = gs[q];
#pragma unroll 1
        for (int m = 0; m < n_g; m++) {
          const cplx a = gs[(m + 1) * gsz + q];
          const double v = cvs[st * n_g + m];
          g.x = fma(v, a.x, g.x);
          g.y = fma(v, a.y, g.y);
        }
        Gb[q] = g;
      }
    } else {
#pragma unroll 1
      for (int q = tid; q < gsz; q += nth) {
        cplx g = sp.g0[q];
#pragma unroll 1
        for (int m = 0; m < n_g; m++) {
          const cplx a = sp.gk[m * gsz + q];
          const double v = cvs[st * n_g + m];
          g.x = fma(v, a.x, g.x);
          g.y = fma(v, a.y, g.y);
        }
        Gb[q] = g;
      }
    }
  };
  const int y_hi = L.ybytes - C;
  auto yaddr = [&](int a) -> int { return GUARD ? a : min(max(a, 0), y_hi); };
  const bool lreal = sp.lreal != 0;
  const unsigned char* lam = smem_diag + L.off_lam;

  // kout = W ys - i (G' Y - Y G'^dag) + collapse terms, G' = G minus its static main diagonal
  auto compute_K = [&](const unsigned char* __restrict__ Yb, const unsigned char* __restrict__ Gb, int st) {
    int oi_ = oi, oj_ = oj, yb_ = ybase, sl_ = SL, sr_ = SR;
    asm volatile("" : "+r"(oi_), "+r"(oj_), "+r"(yb_), "+r"(sl_), "+r"(sr_));
    const int oi = oi_, oj = oj_, ybase = yb_, SL = sl_, SR = sr_;
#pragma unroll
    for (int e = 0; e < EPT; e++) kout[e] = cmul(WREG ? W[WREG ? e : 0] : wst[e * nth + tid], ys[e]);
    if (sp.g0_tab >= 0) {      // time-dependent main diagonal
      const unsigned char* gv = Gb + sp.g0_tab;
      cplx vj[D];
#pragma unroll
      for (int c2 = 0; c2 < D; c2++) vj[c2] = *(const cplx*)(gv + oj + c2 * SR);
#pragma unroll
      for (int a = 0; a < D; a++) {
        const cplx vi = *(const cplx*)(gv + oi + a * SR);
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) cfma(kout[a * D + c2], cmake(vi.y + vj[c2].y, -(vi.x - vj[c2].x)), ys[a * D + c2]);   // -i (v[i] - conj(v[j]))
      }
    }
    // ---- terms of the local mode: sources are this thread's own registers
#pragma unroll
    for (int dq = 0; dq < 2 * (D - 1); dq++) {
      blk_fence();
      const int da = (dq < D - 1) ? dq + 1 : -(dq - (D - 1) + 1);      // +1 .. +(D-1), -1 .. -(D-1)
      const int tabo = bl.gl_tab[da + BLK_MAXD - 1];
      if (tabo >= 0) {
        const unsigned char* gv = Gb + tabo;
#pragma unroll
        for (int a = 0; a < D; a++) {
          if (a + da < 0 || a + da >= D) continue;
          const cplx vi = *(const cplx*)(gv + oi + a * SR);
          const cplx mvi = cmake(vi.y, -vi.x);            // -i v[i]
#pragma unroll
          for (int c2 = 0; c2 < D; c2++) cfma(kout[a * D + c2], mvi, ys[(a + da) * D + c2]);
        }
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) {
          if (c2 + da < 0 || c2 + da >= D) continue;
          const cplx vj = *(const cplx*)(gv + oj + c2 * SR);
          const cplx w = cmake(vj.y, vj.x);               // +i conj(v[j])
#pragma unroll
          for (int a = 0; a < D; a++) cfma(kout[a * D + c2], w, ys[a * D + c2 + da]);
        }
      }
      const int nl = bl.nLl[da + BLK_MAXD - 1];
      for (int t = 0; t < nl; t++) {
        const unsigned char* la = lam + bl.ll_a[da + BLK_MAXD - 1][t] + oi;
        const unsigned char* lb = lam + bl.ll_b[da + BLK_MAXD - 1][t] + oj;
        const int ri = bl.ll_r[da + BLK_MAXD - 1][t];
        const double r = ri < 0 ? 1.0 : cv[st * n_g + ri];
        if (lreal) {
          double wb[D];
#pragma unroll
          for (int c2 = 0; c2 < D; c2++) wb[c2] = (c2 + da < 0 || c2 + da >= D) ? 0.0 : *(const double*)(lb + c2 * SR);
#pragma unroll
          for (int a = 0; a < D; a++) {
            if (a + da < 0 || a + da >= D) continue;
            const double ra = r * (*(const double*)(la + a * SR));
#pragma unroll
            for (int c2 = 0; c2 < D; c2++) {
              if (c2 + da < 0 || c2 + da >= D) continue;
              const double w = ra * wb[c2];
              const cplx yv = ys[(a + da) * D + c2 + da];
              kout[a * D + c2].x = fma(w, yv.x, kout[a * D + c2].x);
              kout[a * D + c2].y = fma(w, yv.y, kout[a * D + c2].y);
            }
          }
        } else {
#pragma unroll
          for (int a = 0; a < D; a++) {
            if (a + da < 0 || a + da >= D) continue;
            const cplx ca = seq::cscale(r, *(const cplx*)(la + a * SR));
#pragma unroll
            for (int c2 = 0; c2 < D; c2++) {
              if (c2 + da < 0 || c2 + da >= D) continue;
              const cplx cb = *(const cplx*)(lb + c2 * SR);
              const cplx w = cmake(ca.x * cb.x + ca.y * cb.y, ca.y * cb.x - ca.x * cb.y);      // lam_a[i] conj(lam_b[j])
              cfma(kout[a * D + c2], w, ys[(a + da) * D + c2 + da]);
            }
          }
        }
      }
    }
    // ---- terms of the other modes: gathered from Y
    blk_fence();
#pragma unroll 1
    for (int t = 0; t < sp.nGo; t++) {
      const unsigned char* gv = Gb + sp.go_tab[t];
      const int offL = sp.go_offL[t] + ybase, offR = sp.go_offR[t] + ybase;
      cplx vj[D];
#pragma unroll
      for (int c2 = 0; c2 < D; c2++) vj[c2] = *(const cplx*)(gv + oj + c2 * SR);
#pragma unroll
      for (int a = 0; a < D; a++) {
        const cplx vi = *(const cplx*)(gv + oi + a * SR);
        const cplx mvi = cmake(vi.y, -vi.x);
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) {
          const cplx ya = *(const cplx*)(Yb + yaddr(offL + a * SL + c2 * SR));
          cfma(kout[a * D + c2], mvi, ya);
          const cplx yv = *(const cplx*)(Yb + yaddr(offR + a * SL + c2 * SR));
          cfma(kout[a * D + c2], cmake(vj[c2].y, vj[c2].x), yv);
        }
      }
    }
    // collapse terms on the own element with a time-dependent rate
#pragma unroll 1
    for (int t = 0; t < sp.nLo; t++) {
      const unsigned char* la = lam + sp.lo_a[t] + oi;
      const unsigned char* lb = lam + sp.lo_b[t] + oj;
      const double r = cv[st * n_g + sp.lo_r[t]];
      cplx cb[D];
#pragma unroll
      for (int c2 = 0; c2 < D; c2++) cb[c2] = *(const cplx*)(lb + c2 * SR);
#pragma unroll
      for (int a = 0; a < D; a++) {
        const cplx ca = seq::cscale(r, *(const cplx*)(la + a * SR));
#pragma unroll
        for (int c2 = 0; c2 < D; c2++)
          cfma(kout[a * D + c2], cmake(ca.x * cb[c2].x + ca.y * cb[c2].y, ca.y * cb[c2].x - ca.x * cb[c2].y), ys[a * D + c2]);
      }
    }
    // collapse terms that gather rho[i + o_a, j + o_b]
#pragma unroll 1
    for (int t = 0; t < sp.nLg; t++) {
      const unsigned char* la = lam + sp.lg_a[t] + oi;
      const unsigned char* lb = lam + sp.lg_b[t] + oj;
      const int off = sp.lg_off[t] + ybase;
      const int ri = sp.lg_r[t];
      const double r = ri < 0 ? 1.0 : cv[st * n_g + ri];
      if (lreal) {
        double wb[D];
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) wb[c2] = *(const double*)(lb + c2 * SR);
#pragma unroll
        for (int a = 0; a < D; a++) {
          const double ra = r * (*(const double*)(la + a * SR));
#pragma unroll
          for (int c2 = 0; c2 < D; c2++) {
            const cplx yv = *(const cplx*)(Yb + yaddr(off + a * SL + c2 * SR));
            const double w = ra * wb[c2];
            kout[a * D + c2].x = fma(w, yv.x, kout[a * D + c2].x);
            kout[a * D + c2].y = fma(w, yv.y, kout[a * D + c2].y);
          }
        }
      } else {
        cplx cb[D];
#pragma unroll
        for (int c2 = 0; c2 < D; c2++) cb[c2] = *(const cplx*)(lb + c2 * SR);
#pragma unroll
        for (int a = 0; a < D; a++) {
          const cplx ca = seq::cscale(r, *(const cplx*)(la + a * SR));
#pragma unroll
          for (int c2 = 0; c2 < D; c2++) {
            const cplx yv = *(const cplx*)(Yb + yaddr(off + a * SL + c2 * SR));
            cfma(kout[a * D + c2], cmake(ca.x * cb[c2].x + ca.y * cb[c2].y, ca.y * cb[c2].x - ca.x * cb[c2].y), yv);
          }
        }
      }
    }
  };
  auto emit = [&](const unsigned char* __restrict__ Yb, int o) {
    // one warp per operator, starting from the LAST warp (warp 0 carries the step controller)
    for (int e = nw - 1 - warp; e < p.n_e; e += nw) {
      cplx acc = cmake(0, 0);
      const int q0 = sp.e_off[e], q1 = sp.e_off[e + 1];
      for (int q = q0 + lane; q < q1; q += 32) {
        const unsigned code = eij[q];
        if (sp.stage_e) {      // resolved at kernel start
          cplx r = ((const cplx*)Yb)[code & 0x7fffffffu];
          if (code >> 31) r.y = -r.y;
          cfma(acc, eval_[q], r);
        } else {
          cfma(acc, eval_[q], ((const cplx*)Yb)[(code >> 16) * ld + (code & 0xffffu)]);   // E_ij rho_ji
        }
      }
      if (q1 - q0 > 1) {          // projectors (the sweeps' e_ops) have a single entry: nothing to reduce
        acc.x = warp_sum(acc.x);
        acc.y = warp_sum(acc.y);
      }
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

  DiagCtlState* cst = (DiagCtlState*)(smem_diag + L.off_ctrl + 128);
  if (tid == 0) {
    StepCtl c0;
    ctl_init(c0, p);
    cst->ctl = c0;
    cst->htry = 0.0; cst->target = 0.0; cst->o = 0; cst->nst = 0; cst->landing = 0; cst->capped = 0; cst->dbg = 0; cst->nstA = 0;
  }
  __syncwarp();
  write_Y(L.off_Y, ys);
  if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 0, 0, 0.0, 0);
  __syncthreads();

  // acc[e] += c * k_j[e] for a stage derivative that lives in a slot (one uniform branch, then immediate offsets)
  auto k_axpy = [&](int j, const double c, cplx (&acc)[EPT]) {
    const int slot = K1REG ? j - 1 : j;
    if (slot < kslots) {
      const cplx* src = ksm + slot * (EPT * nth) + tid;
#pragma unroll
      for (int e = 0; e < EPT; e++) { const cplx kv = src[e * nth]; acc[e].x = fma(c, kv.x, acc[e].x); acc[e].y = fma(c, kv.y, acc[e].y); }
    } else {
      const cplx* src = kgl + (slot - kslots) * (EPT * nth) + tid;
#pragma unroll
      for (int e = 0; e < EPT; e++) { const cplx kv = src[e * nth]; acc[e].x = fma(c, kv.x, acc[e].x); acc[e].y = fma(c, kv.y, acc[e].y); }
    }
  };
  auto kstore = [&](int j, const cplx (&v)[EPT]) {
    const int slot = K1REG ? j - 1 : j;
    if (slot < kslots) {
      cplx* dst = ksm + slot * (EPT * nth) + tid;
#pragma unroll
      for (int e = 0; e < EPT; e++) dst[e * nth] = v[e];
    } else {
      cplx* dst = kgl + (slot - kslots) * (EPT * nth) + tid;
#pragma unroll
      for (int e = 0; e < EPT; e++) dst[e * nth] = v[e];
    }
  };

  // One embedded-RK step.  Returns this thread's share of the weighted squared error.  The arithmetic (order of the
  // multiply-adds of a stage combination and of the error estimate) is the one of every other kernel.
  auto rhs_stage = [&](int st) {      // Y <- ys, G(t) of stage st, kout <- f(ys)
    if (sp.ybufs == 2) yb ^= 1;
    gb ^= 1;
    write_Y(L.off_Y + yb * L.ystride, ys);
    build_Gt((cplx*)(smem_diag + L.off_Gt + gb * L.gtbytes), cv, st);
    PROF(1);
    __syncthreads();
    PROF(2);
    compute_K(smem_diag + L.off_Y + yb * L.ystride, smem_diag + L.off_Gt + gb * L.gtbytes, st);
    if (sp.ybufs == 1) __syncthreads();       // all gathers done before the next stage overwrites Y
  };
  auto run_step = [&](const int pair, const int st0, const double htry) -> double {
    const int S = pair ? 5 : 7;
    if (st0 == 0) {
      // k1 is not known (first step of a trajectory): its evaluation is peeled off the stage loop so that inside the loop
      // k1 and ys have ONE definition each (otherwise the compiler carries two register copies of both through it)
#pragma unroll
      for (int e = 0; e < EPT; e++) ys[e] = y[e];
      rhs_stage(0);
      if (K1REG) {
#pragma unroll
        for (int e = 0; e < EPT; e++) k1[K1REG ? e : 0] = kout[e];
      } else {
        kstore(0, kout);
      }
    }
#pragma unroll 1
    for (int st = 1; st < S; st++) {
      cplx acc[EPT];
      if (K1REG) {      // (a_{st,0} k1 as a product: what fma(a, k, 0) of the other kernels rounds to)
        const double a0 = RK_A[pair][st][0];
#pragma unroll
        for (int e = 0; e < EPT; e++) acc[e] = cmake(a0 * k1[K1REG ? e : 0].x, a0 * k1[K1REG ? e : 0].y);
      } else {
#pragma unroll
        for (int e = 0; e < EPT; e++) acc[e] = cmake(0, 0);
      }
#pragma unroll 1
      for (int j = K1REG ? 1 : 0; j < st; j++) {
        const double aj = RK_A[pair][st][j];
        if (aj != 0.0) k_axpy(j, aj, acc);
      }
#pragma unroll
      for (int e = 0; e < EPT; e++) ys[e] = cmake(fma(htry, acc[e].x, y[e].x), fma(htry, acc[e].y, y[e].y));
      rhs_stage(st);
      if (st < S - 1) kstore(st, kout);
      PROF(3);
    }
    // error estimate: h (sum_{j<S-1} e_j k_j + e_{S-1} k_S), weighted RMS over all N^2 elements
    // (the last stage input ys is the new state; weights 1, or 2 with the mirrored element)
    cplx er[EPT];
    {
      const double cl = RK_E[pair][S - 1];
#pragma unroll
      for (int e = 0; e < EPT; e++) er[e] = cmake(cl * kout[e].x, cl * kout[e].y);
    }
    if (K1REG) {
      const double c0 = RK_E[pair][0];
#pragma unroll
      for (int e = 0; e < EPT; e++) { er[e].x = fma(c0, k1[K1REG ? e : 0].x, er[e].x); er[e].y = fma(c0, k1[K1REG ? e : 0].y, er[e].y); }
    }
#pragma unroll 1
    for (int j = K1REG ? 1 : 0; j < S - 1; j++) {
      const double cj = RK_E[pair][j];
      if (cj != 0.0) k_axpy(j, cj, er);
    }
    double es = 0.0;
    // (threads without an element are masked at the end - their gathers read real data, so their numbers are not zero -
    // because a zero weight would send the division to its slow path.
    // Likewise the square root of an exact zero - most of rho is - takes a slow path: below 1e-280 the floor changes
    // nothing in atol + rtol sqrt(.) as long as atol is not absurdly small)
    const double wgt = isdiag ? 1.0 : 2.0;
    const double m2floor = p.atol >= 1e-100 ? 1e-280 : 0.0;
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const double sc = p.atol + p.rtol * sqrt(fmax(fmax(cabs2(y[e]), cabs2(ys[e])), m2floor));
      const double ex = er[e].x * htry, ey = er[e].y * htry;
      es = fma(ex * ex + ey * ey, wgt / (sc * sc), es);
    }
    return valid ? es : 0.0;
  };

  while (true) {
    const DiagCtrl c = *ctrl;
    if (c.accepted) {
#pragma unroll
      for (int e = 0; e < EPT; e++) y[e] = ys[e];
      if (K1REG) {
#pragma unroll
        for (int e = 0; e < EPT; e++) k1[K1REG ? e : 0] = kout[e];
      } else {
        kstore(0, kout);
      }
      ycur = yb;
    }
    for (int o = c.emit_lo; o < c.emit_hi; o++) emit(smem_diag + L.off_Y + ycur * L.ystride, o);
    PROF(0);
    if (c.cmd) break;
    if (sp.ybufs == 1 && c.emit_hi > c.emit_lo) __syncthreads();     // the next stage overwrites the buffer emit() read
    double es = run_step(c.pair, c.st0, c.htry);
    PROF(4);
    es = warp_sum(es);
    if (lane == 0) red[tog * 32 + warp] = es;
    __syncthreads();
    es = 0.0;
    for (int q = 0; q < nw; q++) es += red[tog * 32 + q];
    tog ^= 1;
    PROF(5);
    if (warp == 0) diag_controller(p, cst, ctrl, cv, tabc, tabr, cscale, 1, (c.pair ? 5 : 7) - c.st0, es, 0);
    PROF(6);
    __syncthreads();
    PROF(7);
  }

  if (valid) {
    cplx* fin = p.final_state + (long long)b * N * N;
    const int bi = oi / C, bj = oj / C, s = bl.s;
#pragma unroll
    for (int a = 0; a < D; a++)
#pragma unroll
      for (int c2 = 0; c2 < D; c2++) {
        const int i = bi + a * s, j = bj + c2 * s;
        fin[i * N + j] = y[a * D + c2];
        if (!isdiag) fin[j * N + i] = cconj(y[a * D + c2]);
      }
  }
  if (tid == 0) {
    p.status[b] = cst->ctl.status;
    p.stats[b * 4 + 0] = cst->ctl.nacc;
    p.stats[b * 4 + 1] = cst->ctl.nrej;
    p.stats[b * 4 + 2] = cst->ctl.nrhs;
    p.stats[b * 4 + 3] = cst->dbg;
  }
#ifdef SEQ_PROFILE_REGIONS
  if (tid == 0) {
    const unsigned long long dc = clock64() - prof_c0;
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0) { seq_prof[13] += dc; seq_prof[15] += g1 - prof_g0; }
    atomicMax(&seq_prof[14], dc);
  }
#endif
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct BlkPlan {
  DiagPlan dp;          // gathered terms, tables, shared-memory layout (dp.ops), table extraction (dp.fill)
  BlkOps bl;
  int threads, maxt, minb;
  bool k1reg, wreg;
  int gslots;           // stage-derivative slots in global memory
  int gathers;          // shared-memory gathers per element and RHS that remain
  int gathers_plain;    // ... of the plain diagonal kernel on the same operators
  long long fma_per_rhs;   // FP64 multiply-adds one right-hand side executes (all threads of a trajectory)
  long long fp64_step_extra;   // FP64 pipe instructions of one 4(3) step outside the right-hand sides: stage combinations (28 per element) and
                               // the error estimate (41 per element, of which ~18 are the IEEE square root and division sequences)
};

// Candidate local modes from the diagonal offsets that occur: (s, D) with D in {3, 2, 4} and N a multiple of s D.
static inline void diagblk_candidates(int N, int n_l, const int* flags, BlkCands* c) {
  const int W = 2 * N - 1;
  c->n = 0;
  auto add = [&](int s, int D) {
    if (s <= 0 || N % (s * D) != 0 || N / D < 2) return;
    for (int q = 0; q < c->n; q++) if (c->s[q] == s && c->D[q] == D) return;
    if (c->n < 8) { c->s[c->n] = s; c->D[c->n] = D; c->n++; }
  };
  const int Ds[2] = {3, 2};
  // offsets of the time-dependent part first (the driven mode), then of the collapse operators, then of G0
  const int order[3] = {1 + n_l, -1, 0};
  for (int pass = 0; pass < 3; pass++)
    for (int o = 1; o < N; o++) {
      bool hit = false;
      if (order[pass] >= 0) hit = flags[(long long)order[pass] * W + o + N - 1] || flags[(long long)order[pass] * W - o + N - 1];
      else for (int l = 0; l < n_l && !hit; l++) hit = flags[(long long)(1 + l) * W + o + N - 1] || flags[(long long)(1 + l) * W - o + N - 1];
      if (hit) for (int D : Ds) add(o, D);
    }
}

// flags: diag_flags_kernel's output; cross: diagblk_cross_kernel's output for candidate (s, D) ([1 + n_l][2N-1]).
static inline bool diagblk_plan(int N, int n_g, int n_c, int n_ctd, int n_ch, int n_t, int nnzE, const int* flags, const int* cross, const int* irreg, int s, int D,
                                int max_smem, int sm_count, BlkPlan* bp) {
  const int n_l = n_c + n_ctd, Wd = 2 * N - 1, c = (int)sizeof(cplx);
  if (N < 2 || N > 0x7fff || D < 2 || D > 3 || N % (s * D) != 0) return false;
  DiagOps& o = bp->dp.ops;
  DiagFill& f = bp->dp.fill;
  BlkOps& bl = bp->bl;
  memset(&o, 0, sizeof(o));
  memset(&f, 0, sizeof(f));
  memset(&bl, 0, sizeof(bl));
  for (int q = 0; q < BLK_ND; q++) bl.gl_tab[q] = -1;
  o.ld = N | 1;
  o.nD = 0; o.g0_tab = -1; o.nGo = 0;
  bl.D = D; bl.s = s; bl.R = N / D;
  bl.SL = s * o.ld * c; bl.SR = s * c;
  auto is_local = [&](int op, int off) { return off % s == 0 && abs(off / s) < D && !cross[(long long)op * Wd + off + N - 1]; };
  int maxoff = 0, gath = 0, gath_plain = 0;
  bool need_mirror = false;      // some gather may read a block in the orientation its owner does not store
  long long fma = 0;      // per outer pair (D x D block)
  const int E2 = D * D;
  fma += 4LL * E2;        // static weight product
  for (int q = 0; q < Wd; q++) {
    if (!flags[q]) continue;
    if (o.nD >= DIAG_MAX_GD) return false;
    const int off = q - (N - 1);
    f.goff[o.nD] = off;
    if (off == 0) {
      if (!flags[(long long)(1 + n_l) * Wd + q]) continue;     // static main diagonal only: folded into the own weight
      o.g0_tab = o.nD * N * c;
      fma += 4LL * E2;
    } else if (is_local(0, off)) {
      const int da = off / s;
      bl.gl_tab[da + BLK_MAXD - 1] = o.nD * N * c;
      fma += 2LL * 4 * (D - abs(da)) * D;
      gath_plain += 2;
    } else {
      o.go_tab[o.nGo] = o.nD * N * c;
      o.go_offL[o.nGo] = off * o.ld * c;
      o.go_offR[o.nGo] = off * c;
      o.nGo++;
      maxoff = abs(off) > maxoff ? abs(off) : maxoff;
      fma += 2LL * 4 * E2;
      gath += 2; gath_plain += 2;
      need_mirror = true;      // (a row or a column alone is shifted: the pair changes its band position)
    }
    o.nD++;
  }
  if (o.nD == 0) { f.goff[0] = 0; o.nD = 1; }
  f.nD = o.nD;
  int nT = 0;
  o.nLs = o.nLo = o.nLg = 0;
  for (int l = 0; l < n_l; l++) {
    const int first = nT;
    for (int q = 0; q < Wd; q++) {
      if (!flags[(long long)(1 + l) * Wd + q]) continue;
      if (nT >= DIAG_MAX_TAB) return false;
      f.t_op[nT] = (short)l; f.t_off[nT] = q - (N - 1);
      nT++;
    }
    if (nT - first > 4) return false;
    const int rate = (l < n_c) ? -1 : (n_ch + (l - n_c));
    for (int a = first; a < nT; a++)
      for (int b2 = first; b2 < nT; b2++) {
        const int oa = f.t_off[a], ob = f.t_off[b2];
        if (oa == 0 && ob == 0) {
          if (o.nLo >= DIAG_MAX_LT || o.nLs >= DIAG_MAX_LT) return false;
          if (rate < 0) { o.ls_a[o.nLs] = a * N * c; o.ls_b[o.nLs] = b2 * N * c; o.nLs++; }
          else { o.lo_a[o.nLo] = a * N * c; o.lo_b[o.nLo] = b2 * N * c; o.lo_r[o.nLo] = rate; o.nLo++; fma += 6LL * E2; }
        } else if (oa == ob && is_local(1 + l, oa) && bl.nLl[oa / s + BLK_MAXD - 1] < BLK_MAXLL) {
          const int q = oa / s + BLK_MAXD - 1, t = bl.nLl[q]++;
          bl.ll_a[q][t] = a * N * c; bl.ll_b[q][t] = b2 * N * c; bl.ll_r[q][t] = rate;
          fma += 3LL * (D - abs(oa / s)) * (D - abs(oa / s));
          gath_plain += 1;
        } else {
          if (o.nLg >= DIAG_MAX_LT) return false;
          o.lg_a[o.nLg] = a * N * c; o.lg_b[o.nLg] = b2 * N * c; o.lg_r[o.nLg] = rate;
          o.lg_off[o.nLg] = (oa * o.ld + ob) * c;
          o.nLg++;
          if (oa != ob || irreg[(long long)(1 + l) * Wd + oa + N - 1]) need_mirror = true;
          maxoff = abs(oa) > maxoff ? abs(oa) : maxoff;
          maxoff = abs(ob) > maxoff ? abs(ob) : maxoff;
          fma += 3LL * E2;
          gath += 1; gath_plain += 1;
        }
      }
  }
  o.nT = nT; f.nT = nT;
  o.nnzE = nnzE;
  bp->dp.maxoff = maxoff;
  bp->gathers = gath; bp->gathers_plain = gath_plain;
  bl.mirror = need_mirror ? 1 : 0;      // (launch_diag also switches it on for stored states and unstaged expectation lists)
  o.n_el = N * (N + 1) / 2;
  // launch shape: one thread per outer pair, lpr lanes per outer row
  const int R = bl.R, Lb = R / 2 + 1;
  bl.lpr = (Lb + 7) / 8 * 8;
  bp->threads = (R * bl.lpr + 31) / 32 * 32;
  if (bp->threads > 512) return false;
  bp->fma_per_rhs = fma * ((long long)R * (R + 1) / 2);
  bp->fp64_step_extra = 69LL * D * D * ((long long)R * (R + 1) / 2);
  o.nct = bp->threads;
  const int E = D * D;
  const int guard_el = maxoff * o.ld + maxoff + 1;
  // launch shape by thread count: small CTAs run two per SM (255 registers each), the largest ones keep nothing but the
  // state, the stage input and the accumulator in registers (128 registers at 512 threads)
  if (bp->threads <= 128) {
    bp->maxt = 128; bp->minb = 2; bp->k1reg = true; bp->wreg = true;
    const char* mb = getenv("SEQ_BLK_MINB");      // experiments: 3 CTAs per SM (168 registers, more slots in L2)
    if (mb && atoi(mb) == 3 && D == 3) bp->minb = 3;
  }
  else if (bp->threads <= 256) { bp->maxt = 256; bp->minb = 1; bp->k1reg = true; bp->wreg = true; }
  else { bp->maxt = 512; bp->minb = 1; bp->k1reg = false; bp->wreg = false; }
  const int slot = E * bp->maxt * c;           // one thread-private stage vector (the kernel runs with maxt threads)
  const int nslots = bp->k1reg ? 5 : 6;        // k2 .. k6, or k1 .. k6
  const int rk43 = nslots - 2;                 // what the 4(3) pair needs (k5 / k6 belong to the 5(4) pair)
  const int per_cta = (233472 / bp->minb) - 1024 < max_smem ? (233472 / bp->minb) - 1024 : max_smem;      // 228 KB per SM, 1 KB per CTA reserved
  bl.wslot = bp->wreg ? 0 : 2;
  // in order of value: the 4(3) pair's derivatives on chip, then a second Y buffer (one barrier per RHS instead of two),
  // then k5 / k6
  const int tries[8][2] = {{nslots, 2}, {rk43, 2}, {nslots, 1}, {rk43, 1}, {2, 1}, {1, 1}, {0, 1}, {0, 1}};
  bool ok = false;
  for (int t = 0; t < 7 && !ok; t++) {
    bl.kslots = tries[t][0];
    o.ybufs = tries[t][1];
    o.wsm = bl.kslots * slot;
    for (int g = 1; g >= 0 && !ok; g--) {
      bp->dp.guard = g != 0;
      o.guard = g ? guard_el : 0;
      o.stage_g = o.stage_e = o.stage_tab = 0;
      ok = diag_layout(N, n_g, n_t, o).total <= per_cta;
    }
  }
  if (!ok) return false;
  if (!bp->wreg) {      // static weights on chip if they still fit
    o.wsm += slot;
    if (diag_layout(N, n_g, n_t, o).total <= per_cta) bl.wslot = 1; else o.wsm -= slot;
  }
  o.stage_g = 1;
  if (diag_layout(N, n_g, n_t, o).total > per_cta) o.stage_g = 0;
  o.stage_e = (nnzE > 0 && nnzE <= DIAG_E_SMEM) ? 1 : 0;
  if (o.stage_e && diag_layout(N, n_g, n_t, o).total > per_cta) o.stage_e = 0;
  o.stage_tab = (long long)n_g * n_t * 16 < (1 << 20) ? 1 : 0;
  if (o.stage_tab && diag_layout(N, n_g, n_t, o).total > per_cta) o.stage_tab = 0;
  bp->gslots = nslots - bl.kslots;
  bp->dp.smem = (size_t)diag_layout(N, n_g, n_t, o).total;
  bp->dp.kreg = true; bp->dp.cl = 1; bp->dp.cw = false; bp->dp.ept = E; bp->dp.threads = bp->threads;
  (void)sm_count;
  return true;
}

static inline size_t diagblk_ws_elems(const BlkPlan& bp) { return (size_t)bp.gslots * bp.bl.D * bp.bl.D * bp.maxt; }
static inline size_t diagblk_ws_shared_elems(const BlkPlan& bp) { return bp.bl.wslot == 2 ? (size_t)bp.bl.D * bp.bl.D * bp.maxt : 0; }

#if defined(SEQ_BLK_EXTERN)
// split build: the kernel instantiations live in their own translation unit (csrc/diagblk_unit.cu, compiled in parallel
// with seq_engine.cu inside a renamed namespace) behind one C entry point
extern "C" int seq_blk_launch_ext(const void* p, const void* bp, void* stream);
static inline int launch_diagblk_kernel(const SolveParams& p, const BlkPlan& bp, cudaStream_t st) { return seq_blk_launch_ext(&p, &bp, (void*)st); }
#elif !defined(SEQ_DIAG_NO_LAUNCH)
template <int D, bool K1REG, bool WREG, bool GUARD, int MAXT, int MINB>
static int launch_diagblk_t(const SolveParams& p, const BlkPlan& bp, cudaStream_t st) {
  auto kern = solve_diagblk_kernel<D, K1REG, WREG, GUARD, MAXT, MINB>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.dp.smem) != cudaSuccess) return SEQ_ERR_CUDA;
  kern<<<(unsigned)p.B, (unsigned)MAXT, bp.dp.smem, st>>>(p, bp.dp.ops, bp.bl);
  return cudaGetLastError() == cudaSuccess ? SEQ_OK : SEQ_ERR_CUDA;
}

template <int D>
static inline int launch_diagblk_d(const SolveParams& p, const BlkPlan& bp, cudaStream_t st) {
  const bool g = bp.dp.guard;
  if (bp.maxt == 128 && bp.minb == 3 && D == 3) return g ? launch_diagblk_t<3, true, true, true, 128, 3>(p, bp, st) : launch_diagblk_t<3, true, true, false, 128, 3>(p, bp, st);
  if (bp.maxt == 128) return g ? launch_diagblk_t<D, true, true, true, 128, 2>(p, bp, st) : launch_diagblk_t<D, true, true, false, 128, 2>(p, bp, st);
  if (bp.maxt == 256) return g ? launch_diagblk_t<D, true, true, true, 256, 1>(p, bp, st) : launch_diagblk_t<D, true, true, false, 256, 1>(p, bp, st);
  return g ? launch_diagblk_t<D, false, false, true, 512, 1>(p, bp, st) : launch_diagblk_t<D, false, false, false, 512, 1>(p, bp, st);
}

static inline int launch_diagblk_kernel(const SolveParams& p, const BlkPlan& bp, cudaStream_t st) {
  if (bp.bl.D == 3) return launch_diagblk_d<3>(p, bp, st);
  if (bp.bl.D == 2) return launch_diagblk_d<2>(p, bp, st);
  return SEQ_ERR_UNSUPPORTED;
}
#endif

}  // namespace seq
