// SEQ_METHOD_GENERIC: one CTA per trajectory, FP64 FMA, all state in per-trajectory global
// scratch (L1/L2 resident).  Covers any N, density matrices and kets, and is the on-device
// cross-check for the tensor-core kernels.  The whole time span is integrated inside one launch:
// Dormand-Prince 5(4) with FSAL, per-trajectory (CTA-uniform) adaptive step control, fused
// spline evaluation, expectation values Tr(E rho) reduced with warp shuffles at output samples.
#pragma once
#include "common.cuh"

namespace seq {

struct GenericCtx {
  const SolveParams* p;
  int b, N, NN;          // NN = N*N (dm) or N (ket): number of state elements
  const cplx* G0;
  const double2* tabc;   // this trajectory's coefficient tables
  const double2* tabr;   // this trajectory's rate tables
  const double* cscale;  // this trajectory's coeff_scale row (or null)
  cplx* G;               // [N,N] scratch: G(t)
  cplx* T;               // [N,N] scratch
  double* cval;          // smem: n_g spline values
  double* red;           // smem: reduction scratch (33 doubles)
};

// G(t) = G0 + sum_m g_m(t) Gk[m]  -> ctx.G ; leaves coefficient values in ctx.cval
__device__ __forceinline__ void build_G(const GenericCtx& c, double t) {
  const SolveParams& p = *c.p;
  for (int m = threadIdx.x; m < p.n_g; m += blockDim.x) {      // (up to 64 terms on as few as 32 threads)
    double v;
    if (m < p.n_ch) {
      v = spline_eval(c.tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, t);
      if (c.cscale) v *= c.cscale[m];
    } else {
      v = spline_eval(c.tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, t);
    }
    c.cval[m] = v;
  }
  __syncthreads();
  const int N2 = c.N * c.N;
  for (int idx = threadIdx.x; idx < N2; idx += blockDim.x) {
    cplx g = c.G0[idx];
    for (int m = 0; m < p.n_g; m++) {
      cplx a = p.Gk[(long long)m * N2 + idx];
      double v = c.cval[m];
      g.x = fma(v, a.x, g.x);
      g.y = fma(v, a.y, g.y);
    }
    c.G[idx] = g;
  }
  __syncthreads();
}

// K = -i (G Y) + h.c. + sum_j r_j L_j Y L_j^dag        (density matrix)
__device__ void rhs_dm(const GenericCtx& c, double t, const cplx* __restrict__ Y, cplx* __restrict__ K) {
  const SolveParams& p = *c.p;
  const int N = c.N, N2 = N * N;
  build_G(c, t);
  // T = G Y
  for (int idx = threadIdx.x; idx < N2; idx += blockDim.x) {
    int i = idx / N, j = idx - i * N;
    cplx acc = cmake(0, 0);
    for (int k = 0; k < N; k++) cfma(acc, c.G[i * N + k], Y[k * N + j]);
    c.T[idx] = acc;
  }
  __syncthreads();
  // K = -i T + (-i T)^dag :  K_ij = -i T_ij + i conj(T_ji)
  for (int idx = threadIdx.x; idx < N2; idx += blockDim.x) {
    int i = idx / N, j = idx - i * N;
    cplx a = c.T[idx], bt = c.T[j * N + i];
    // -i*(ar + i ai) = ai - i ar ;  i*conj(b) = i*(br - i bi) = bi + i br
    K[idx] = cmake(a.y + bt.y, -a.x + bt.x);
  }
  __syncthreads();
  const int nL = p.n_c + p.n_ctd;
  for (int l = 0; l < nL; l++) {
    const cplx* L = p.L + (long long)l * N2;
    double r = (l < p.n_c) ? 1.0 : c.cval[p.n_ch + (l - p.n_c)];
    // T = L Y
    for (int idx = threadIdx.x; idx < N2; idx += blockDim.x) {
      int i = idx / N, j = idx - i * N;
      cplx acc = cmake(0, 0);
      for (int k = 0; k < N; k++) cfma(acc, L[i * N + k], Y[k * N + j]);
      c.T[idx] = acc;
    }
    __syncthreads();
    // K += r * T L^dag : (T L^dag)_ij = sum_k T_ik conj(L_jk)
    for (int idx = threadIdx.x; idx < N2; idx += blockDim.x) {
      int i = idx / N, j = idx - i * N;
      cplx acc = cmake(0, 0);
      for (int k = 0; k < N; k++) cfma_conj(acc, c.T[i * N + k], L[j * N + k]);
      cplx kv = K[idx];
      kv.x = fma(r, acc.x, kv.x);
      kv.y = fma(r, acc.y, kv.y);
      K[idx] = kv;
    }
    __syncthreads();
  }
}

// K = -i G psi        (ket)
__device__ void rhs_ket(const GenericCtx& c, double t, const cplx* __restrict__ Y, cplx* __restrict__ K) {
  const int N = c.N;
  build_G(c, t);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    cplx acc = cmake(0, 0);
    for (int k = 0; k < N; k++) cfma(acc, c.G[i * N + k], Y[k]);
    K[i] = cmake(acc.y, -acc.x);
  }
  __syncthreads();
}

template <bool KET>
__device__ __forceinline__ void rhs(const GenericCtx& c, double t, const cplx* Y, cplx* K) {
  if (KET) rhs_ket(c, t, Y, K); else rhs_dm(c, t, Y, K);
}

// expectation values + optional state store at output slot `o`
template <bool KET>
__device__ void emit_output(const GenericCtx& c, int o, cplx* y) {
  const SolveParams& p = *c.p;
  const int N = c.N;
  if (KET && (p.flags & SEQ_FLAG_NORMALIZE)) {
    double s = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += cabs2(y[i]);
    s = block_sum(s, c.red);
    double inv = 1.0 / sqrt(s);
    for (int i = threadIdx.x; i < N; i += blockDim.x) y[i] = cscale(inv, y[i]);
    __syncthreads();
  }
  for (int e = 0; e < p.n_e; e++) {
    const cplx* E = p.E + (long long)e * N * N;
    cplx acc = cmake(0, 0);
    if (KET) {
      // <psi|E|psi> = sum_i conj(psi_i) (E psi)_i
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        cplx row = cmake(0, 0);
        for (int k = 0; k < N; k++) cfma(row, E[i * N + k], y[k]);
        cfma_conj(acc, row, y[i]);
      }
    } else {
      // Tr(E rho) = sum_ij E_ij rho_ji
      for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
        int i = idx / N, j = idx - i * N;
        cfma(acc, E[idx], y[j * N + i]);
      }
    }
    acc = block_sum_c(acc, c.red);
    if (threadIdx.x == 0) {
      long long off = ((long long)c.b * p.n_e + e) * p.n_out + o;
      if (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ((cplx*)p.expect)[off] = acc;
      else ((double*)p.expect)[off] = acc.x;
    }
  }
  if (p.flags & SEQ_FLAG_STORE_STATES) {
    cplx* dst = p.states + ((long long)c.b * p.n_out + o) * c.NN;
    for (int i = threadIdx.x; i < c.NN; i += blockDim.x) dst[i] = y[i];
  }
}

template <bool KET>
__global__ void solve_generic_kernel(const __grid_constant__ SolveParams p) {
  extern __shared__ double smem[];
  GenericCtx c;
  c.p = &p;
  c.b = blockIdx.x;
  c.N = p.N;
  c.NN = KET ? p.N : p.N * p.N;
  const int N2 = p.N * p.N;
  c.G0 = p.G0 + (long long)c.b * p.G0_bstride;
  c.tabc = p.tab + (long long)c.b * p.tab_bstride_c;
  c.tabr = p.tab_rate + (long long)c.b * p.tab_bstride_r;
  c.cscale = p.coeff_scale ? p.coeff_scale + (long long)c.b * p.n_ch : nullptr;
  c.red = smem;
  c.cval = smem + 34;
  cplx* ws = p.ws + (long long)c.b * p.ws_stride;
  c.G = ws; ws += N2;
  c.T = ws; ws += N2;
  cplx* y = ws; ws += c.NN;
  cplx* ys = ws; ws += c.NN;
  cplx* kk[7];
#pragma unroll
  for (int j = 0; j < 7; j++) { kk[j] = ws; ws += c.NN; }
  const int NN = c.NN;
  const int tid = threadIdx.x, nth = blockDim.x;

  const cplx* y0 = p.state0 + (long long)c.b * NN;
  for (int i = tid; i < NN; i += nth) y[i] = y0[i];
  __syncthreads();

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
        for (int st = ctl.have_k1 ? 1 : 0; st < S; st++) {
          // stage input ys = y + h sum_j a_j k_j   (st == S-1: ys = y_new)
          for (int i = tid; i < NN; i += nth) {
            double ax = 0.0, ay = 0.0;
            for (int j = 0; j < st; j++) {
              double aj = RK_A[pair][st][j];
              cplx kv = kk[j][i];
              ax = fma(aj, kv.x, ax);
              ay = fma(aj, kv.y, ay);
            }
            cplx v = y[i];
            ys[i] = cmake(fma(htry, ax, v.x), fma(htry, ay, v.y));
          }
          __syncthreads();
          rhs<KET>(c, ctl.t + RK_C[pair][st] * htry, ys, kk[st]);
          ctl.nrhs++;
        }
        ctl.have_k1 = true;
        // error estimate, RMS norm scaled by atol + rtol*max(|y|,|y_new|)  (ZVODE's weighted RMS norm)
        double es = 0.0;
        for (int i = tid; i < NN; i += nth) {
          double ex = 0.0, ey = 0.0;
          for (int j = 0; j < S; j++) {
            double ej = RK_E[pair][j];
            cplx kv = kk[j][i];
            ex = fma(ej, kv.x, ex);
            ey = fma(ej, kv.y, ey);
          }
          ex *= htry; ey *= htry;
          double sc = p.atol + p.rtol * sqrt(fmax(cabs2(y[i]), cabs2(ys[i])));
          es += (ex * ex + ey * ey) / (sc * sc);
        }
        es = block_sum(es, c.red);
        double err = sqrt(es / (double)NN);
        if (ctl_update(ctl, p, err, htry, landing, capped, t_target, nst)) {
          cplx* tmp = y; y = ys; ys = tmp;
          tmp = kk[0]; kk[0] = kk[S - 1]; kk[S - 1] = tmp;
        }
      }
      if (ctl.status != SEQ_STATUS_OK) break;
    }
    if (KET && (p.flags & SEQ_FLAG_NORMALIZE)) ctl.have_k1 = false;  // state is rescaled below
    emit_output<KET>(c, o, y);
    __syncthreads();
  }
  cplx* fin = p.final_state + (long long)c.b * NN;
  for (int i = tid; i < NN; i += nth) fin[i] = y[i];
  if (tid == 0) {
    p.status[c.b] = ctl.status;
    p.stats[c.b * 4 + 0] = ctl.nacc;
    p.stats[c.b * 4 + 1] = ctl.nrej;
    p.stats[c.b * 4 + 2] = ctl.nrhs;
    p.stats[c.b * 4 + 3] = 0;
  }
}

}  // namespace seq
