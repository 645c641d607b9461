// SEQ_METHOD_SMALL: density matrices with N <= 6 (C1: transmon(3), C3: transmon(4)).
//
// These shapes are far below one DMMA tile, so the design is the opposite of the tensor-core
// kernels: ONE THREAD integrates one trajectory, everything in registers / thread-private local
// memory, no barriers, no shuffles.  Per-lane step sizes differ, control flow does not: every lane
// of a warp executes the same "attempt one RK step" instruction stream and accept/reject is a
// select, so divergent step sizes never serialise the batch (a warp only waits for the lane that
// needs the most attempts inside one output interval).
//
// Arithmetic: a Hermitian rho has N^2 REAL degrees of freedom.  With
//     x[i*N+i] = rho_ii,   x[i*N+j] = Re rho_ij,   x[j*N+i] = Im rho_ij     (i < j)
// the Lindblad right-hand side is a real linear map  x' = (M_0 + sum_m c_m(t) M_m) x  with
// (1 + n_g) real D x D matrices (D = N^2) shared by the whole batch.  They are built once per solve
// by small_superop_kernel from the same G0 / Gk / L operators the other kernels use (so the
// equations are identical: -i(G rho - rho G^dag) + sum_l L_l rho L_l^dag), live in shared memory
// column-major, and are read with broadcast LDS.128 (all lanes read the same address).  One RHS is
// (1 + n_g) D^2 real FMAs: 243 at C1 and 768 at C3, against 8 N^3 (1 + 2 n_c) / 2 = 540 / 1280 in
// complex-product form.
#pragma once
#include "common.cuh"

namespace seq {

#define SMALL_MAX_N 6

template <int N> struct SmallCfg {
  static constexpr int D = N * N;
  static constexpr int DP = (D + 1) & ~1;   // column stride (even, for 16-byte loads)
};

struct SmallExtra {
  const double* M;   // [1+n_g][D][DP]: M[m][j][i] = d x'_i / d x_j of term m
  const double* W;   // [n_e][2][D]: Tr(E rho) = (W[e][0] + i W[e][1]) . x
};

static inline size_t small_smem_bytes(int N, int n_g, int n_e) {
  size_t D = (size_t)N * N, DP = (D + 1) & ~(size_t)1;
  return ((1 + (size_t)n_g) * D * DP + (size_t)n_e * 2 * D + 2) * sizeof(double);
}
static inline bool small_supported(int N, int n_g, int n_e) {
  return N >= 1 && N <= SMALL_MAX_N && small_smem_bytes(N, n_g, n_e) <= 96 * 1024;
}
static inline size_t small_pack_doubles(int N, int n_g, int n_e) { return small_smem_bytes(N, n_g, n_e) / sizeof(double) + 8; }

// ---- prep: real superoperators and expectation weights.  One thread per (term, column); the
// matrices are tiny (N <= 6) so everything is naive loops over thread-local arrays.
__global__ void small_superop_kernel(SolveParams p, const cplx* __restrict__ G0, const cplx* __restrict__ Gk,
                                     double* __restrict__ M, double* __restrict__ W) {
  const int N = p.N, D = N * N, DP = (D + 1) & ~1;
  const int m = blockIdx.x;        // term
  const int c = threadIdx.x;       // column = parameter index (a, b)
  if (m <= p.n_g && c < D) {
    const int a = c / N, b = c - a * N;
    cplx Bm[SMALL_MAX_N * SMALL_MAX_N], R[SMALL_MAX_N * SMALL_MAX_N], T[SMALL_MAX_N * SMALL_MAX_N];
    for (int i = 0; i < D; i++) { Bm[i] = cmake(0, 0); R[i] = cmake(0, 0); }
    if (a == b) Bm[a * N + a] = cmake(1, 0);
    else if (a < b) { Bm[a * N + b] = cmake(1, 0); Bm[b * N + a] = cmake(1, 0); }       // Re rho_ab = 1
    else { Bm[b * N + a] = cmake(0, 1); Bm[a * N + b] = cmake(0, -1); }               // Im rho_ba = 1
    const cplx* G = (m == 0) ? G0 : Gk + (long long)(m - 1) * D;
    // A = -i G B ;  R = A + A^dag
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        cplx acc = cmake(0, 0);
        for (int k = 0; k < N; k++) cfma(acc, G[i * N + k], Bm[k * N + j]);
        T[i * N + j] = cmake(acc.y, -acc.x);
      }
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) R[i * N + j] = cadd(T[i * N + j], cconj(T[j * N + i]));
    // + sum_l L B L^dag over the collapse operators that belong to this term
    int l0 = 0, l1 = 0;
    if (m == 0) { l0 = 0; l1 = p.n_c; }
    else if (m - 1 >= p.n_ch) { l0 = p.n_c + (m - 1 - p.n_ch); l1 = l0 + 1; }
    for (int l = l0; l < l1; l++) {
      const cplx* L = p.L + (long long)l * D;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
          cplx acc = cmake(0, 0);
          for (int k = 0; k < N; k++) cfma(acc, L[i * N + k], Bm[k * N + j]);
          T[i * N + j] = acc;
        }
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
          cplx acc = R[i * N + j];
          for (int k = 0; k < N; k++) cfma_conj(acc, T[i * N + k], L[j * N + k]);
          R[i * N + j] = acc;
        }
    }
    double* col = M + ((long long)m * D + c) * DP;
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        double v = (i == j) ? R[i * N + i].x : ((i < j) ? R[i * N + j].x : R[j * N + i].y);
        col[i * N + j] = v;
      }
    if (DP > D) col[D] = 0.0;
  }
  // expectation weights: one extra block
  if (m == p.n_g + 1) {
    for (int e = 0; e < p.n_e; e++) {
      const cplx* E = p.E + (long long)e * D;
      if (c < D) {
        const int i = c / N, j = c - (c / N) * N;
        cplx w;
        if (i == j) w = E[c];
        else if (i < j) w = cadd(E[i * N + j], E[j * N + i]);            // weight of Re rho_ij
        else { cplx d = csub(E[i * N + j], E[j * N + i]); w = cmake(-d.y, d.x); }  // i (E_ij - E_ji), (i,j) = (row > col): weight of Im rho_ji
        W[((long long)e * 2 + 0) * D + c] = w.x;
        W[((long long)e * 2 + 1) * D + c] = w.y;
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(64) solve_small_kernel(const __grid_constant__ SolveParams p, const __grid_constant__ SmallExtra x) {
  constexpr int D = SmallCfg<N>::D, DP = SmallCfg<N>::DP;
  extern __shared__ __align__(16) double smem_small[];
  double* Ms = smem_small;
  const int nM = (1 + p.n_g) * D * DP;
  double* Ws = Ms + nM;
  for (int i = threadIdx.x; i < nM; i += blockDim.x) Ms[i] = x.M[i];
  for (int i = threadIdx.x; i < p.n_e * 2 * D; i += blockDim.x) Ws[i] = x.W[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;

  const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
  const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
  const double* cscale = p.coeff_scale ? p.coeff_scale + (long long)b * p.n_ch : nullptr;

  const double inv_dt = 1.0 / p.dt;
  // out = (M_0 + sum_m c_m(t) M_m) xin
  auto rhs = [&](double t, const double (&xin)[D], double (&out)[D]) {
#pragma unroll
    for (int i = 0; i < D; i++) out[i] = 0.0;
    for (int m = 0; m <= p.n_g; m++) {
      double cm = 1.0;
      if (m > 0) {
        const int g = m - 1;
        if (g < p.n_ch) {
          cm = spline_eval_inv(tabc + (long long)g * p.n_t, p.n_t, p.t0, inv_dt, t);
          if (cscale) cm *= cscale[g];
        } else {
          cm = spline_eval_inv(tabr + (long long)(g - p.n_ch) * p.n_t, p.n_t, p.t0, inv_dt, t);
        }
      }
      const double* Mm = Ms + m * D * DP;
#pragma unroll
      for (int j = 0; j < D; j++) {
        const double xs = cm * xin[j];
#pragma unroll
        for (int i = 0; i < DP; i += 2) {
          const double2 mm = *reinterpret_cast<const double2*>(Mm + j * DP + i);
          out[i] = fma(mm.x, xs, out[i]);
          if (i + 1 < D) out[i + 1] = fma(mm.y, xs, out[i + 1]);
        }
      }
    }
  };

  // rho0 -> x
  double y[D];
  {
    const cplx* y0 = p.state0 + (long long)b * D;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = i; j < N; j++) {
        cplx v = y0[i * N + j];
        y[i * N + j] = v.x;
        if (j > i) y[j * N + i] = v.y;
      }
  }
  auto write_state = [&](cplx* dst, const double (&v)[D]) {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (i == j) dst[i * N + j] = cmake(v[i * N + i], 0.0);
        else if (i < j) dst[i * N + j] = cmake(v[i * N + j], v[j * N + i]);
        else dst[i * N + j] = cmake(v[j * N + i], -v[i * N + j]);
      }
  };

  double k[7][D];   // thread-private stage derivatives (local memory, L1 resident)
  StepCtl ctl;
  ctl_init(ctl, p);
  const double atol = p.atol, rtol = p.rtol;

  for (int o = 0; o < p.n_out; o++) {
    if (o > 0) {
      const double t_target = p.t0 + p.dt * (double)p.out_idx[o];
      int nst = 0;
      double htry;
      bool landing, capped;
      while (ctl.status == SEQ_STATUS_OK && ctl_next(ctl, p, t_target, htry, landing, capped)) {
        const int pair = ctl.pair, S = RK_STAGES[pair];
        double ys[D];
        for (int st = ctl.have_k1 ? 1 : 0; st < S; st++) {
          double acc[D];
#pragma unroll
          for (int i = 0; i < D; i++) acc[i] = 0.0;
          for (int j = 0; j < st; j++) {
            const double aj = RK_A[pair][st][j];
#pragma unroll
            for (int i = 0; i < D; i++) acc[i] = fma(aj, k[j][i], acc[i]);
          }
#pragma unroll
          for (int i = 0; i < D; i++) ys[i] = fma(htry, acc[i], y[i]);
          double kout[D];
          rhs(ctl.t + RK_C[pair][st] * htry, ys, kout);
#pragma unroll
          for (int i = 0; i < D; i++) k[st][i] = kout[i];
          ctl.nrhs++;
        }
        ctl.have_k1 = true;
        // error estimate: the weighted RMS norm over the N*N complex elements of rho, written in
        // the real parametrisation (an off-diagonal pair (i,j),(j,i) shares one modulus)
        double ev[D];
#pragma unroll
        for (int i = 0; i < D; i++) ev[i] = 0.0;
        for (int j = 0; j < S; j++) {
          const double ej = RK_E[pair][j];
#pragma unroll
          for (int i = 0; i < D; i++) ev[i] = fma(ej, k[j][i], ev[i]);
        }
        double es = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
          for (int j = i; j < N; j++) {
            if (i == j) {
              const double sc = atol + rtol * fmax(fabs(y[i * N + i]), fabs(ys[i * N + i]));
              const double e = htry * ev[i * N + i];
              es += (e * e) / (sc * sc);
            } else {
              const double m0 = y[i * N + j] * y[i * N + j] + y[j * N + i] * y[j * N + i];
              const double m1 = ys[i * N + j] * ys[i * N + j] + ys[j * N + i] * ys[j * N + i];
              const double sc = atol + rtol * sqrt(fmax(m0, m1));
              const double er = htry * ev[i * N + j], ei = htry * ev[j * N + i];
              es += 2.0 * (er * er + ei * ei) / (sc * sc);
            }
          }
        const double err = sqrt(es / (double)D);
        if (ctl_update<true>(ctl, p, err, htry, landing, capped, t_target, nst)) {
#pragma unroll
          for (int i = 0; i < D; i++) { y[i] = ys[i]; k[0][i] = k[S - 1][i]; }
        }
      }
      if (ctl.status != SEQ_STATUS_OK) break;
    }
    for (int e = 0; e < p.n_e; e++) {
      const double* wr = Ws + (e * 2 + 0) * D;
      const double* wi = wr + D;
      double ar = 0.0, ai = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) { ar = fma(wr[i], y[i], ar); ai = fma(wi[i], y[i], ai); }
      const long long off = ((long long)b * p.n_e + e) * p.n_out + o;
      if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = cmake(ar, ai);
      else ((double*)p.expect)[off] = ar;
    }
    if (p.flags & SEQ_FLAG_STORE_STATES) write_state(p.states + ((long long)b * p.n_out + o) * D, y);
  }
  write_state(p.final_state + (long long)b * D, y);
  p.status[b] = ctl.status;
  p.stats[b * 4 + 0] = ctl.nacc;
  p.stats[b * 4 + 1] = ctl.nrej;
  p.stats[b * 4 + 2] = ctl.nrhs;
  p.stats[b * 4 + 3] = 0;
}

template <int N>
static int launch_small_n(const SolveParams& p, const SmallExtra& x, int sms, cudaStream_t st) {
  const size_t smem = small_smem_bytes(N, p.n_g, p.n_e);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(solve_small_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SEQ_ERR_CUDA;
  }
  // small CTAs so that the SMs are evenly loaded: one warp per CTA until every SM has work for two
  const int threads = (p.B > sms * 64) ? 64 : 32;
  const int blocks = (p.B + threads - 1) / threads;
  solve_small_kernel<N><<<blocks, threads, smem, st>>>(p, x);
  return SEQ_OK;
}

static inline int launch_small(const SolveParams& p, const cplx* G0, const cplx* Gk, double* pack, int sms, cudaStream_t st,
                               int* launches) {
  const int N = p.N, D = N * N, DP = (D + 1) & ~1;
  SmallExtra x;
  double* M = pack;
  double* W = M + (size_t)(1 + p.n_g) * D * DP;
  small_superop_kernel<<<p.n_g + 2, 64, 0, st>>>(p, G0, Gk, M, W);
  (*launches)++;
  x.M = M; x.W = W;
  int rc;
  switch (N) {
    case 1: rc = launch_small_n<1>(p, x, sms, st); break;
    case 2: rc = launch_small_n<2>(p, x, sms, st); break;
    case 3: rc = launch_small_n<3>(p, x, sms, st); break;
    case 4: rc = launch_small_n<4>(p, x, sms, st); break;
    case 5: rc = launch_small_n<5>(p, x, sms, st); break;
    case 6: rc = launch_small_n<6>(p, x, sms, st); break;
    default: return SEQ_ERR_UNSUPPORTED;
  }
  if (rc == SEQ_OK) (*launches)++;
  return rc;
}

}  // namespace seq
