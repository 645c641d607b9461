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
  if ((int)threadIdx.x < p.n_g) {
    int m = threadIdx.x;
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
  cplx* k1 = ws; ws += c.NN;
  cplx* k2 = ws; ws += c.NN;
  cplx* k3 = ws; ws += c.NN;
  cplx* k4 = ws; ws += c.NN;
  cplx* k5 = ws; ws += c.NN;
  cplx* k6 = ws; ws += c.NN;
  cplx* k7 = ws; ws += c.NN;
  const int NN = c.NN;
  const int tid = threadIdx.x, nth = blockDim.x;

  const cplx* y0 = p.state0 + (long long)c.b * NN;
  for (int i = tid; i < NN; i += nth) y[i] = y0[i];
  __syncthreads();

  int status = SEQ_STATUS_OK, nacc = 0, nrej = 0, nrhs = 0;
  double t = p.t0 + p.dt * (double)p.out_idx[0];
  double h = p.first_step > 0 ? p.first_step : p.dt;
  if (p.max_step > 0 && h > p.max_step) h = p.max_step;
  bool have_k1 = false, prev_rejected = false;

  for (int o = 0; o < p.n_out && status == SEQ_STATUS_OK; o++) {
    if (o > 0) {
      const double t_target = p.t0 + p.dt * (double)p.out_idx[o];
      int nst = 0;
      while (true) {
        double rem = t_target - t;
        if (rem <= 1e-13 * fmax(1.0, fabs(t_target))) break;
        double htry = h;
        if (p.max_step > 0 && htry > p.max_step) htry = p.max_step;
        double q = ceil(rem / htry - 1e-9);
        if (q < 1.0) q = 1.0;
        htry = rem / q;
        const bool landing = (q == 1.0);
        if (!have_k1) { rhs<KET>(c, t, y, k1); nrhs++; have_k1 = true; }
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i];
          ys[i] = cmake(fma(htry * DP_A21, a.x, v.x), fma(htry * DP_A21, a.y, v.y));
        }
        __syncthreads();
        rhs<KET>(c, t + DP_C2 * htry, ys, k2);
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i], b2 = k2[i];
          ys[i] = cmake(v.x + htry * (DP_A31 * a.x + DP_A32 * b2.x), v.y + htry * (DP_A31 * a.y + DP_A32 * b2.y));
        }
        __syncthreads();
        rhs<KET>(c, t + DP_C3 * htry, ys, k3);
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i], b2 = k2[i], b3 = k3[i];
          ys[i] = cmake(v.x + htry * (DP_A41 * a.x + DP_A42 * b2.x + DP_A43 * b3.x),
                        v.y + htry * (DP_A41 * a.y + DP_A42 * b2.y + DP_A43 * b3.y));
        }
        __syncthreads();
        rhs<KET>(c, t + DP_C4 * htry, ys, k4);
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i], b2 = k2[i], b3 = k3[i], b4 = k4[i];
          ys[i] = cmake(v.x + htry * (DP_A51 * a.x + DP_A52 * b2.x + DP_A53 * b3.x + DP_A54 * b4.x),
                        v.y + htry * (DP_A51 * a.y + DP_A52 * b2.y + DP_A53 * b3.y + DP_A54 * b4.y));
        }
        __syncthreads();
        rhs<KET>(c, t + DP_C5 * htry, ys, k5);
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i], b2 = k2[i], b3 = k3[i], b4 = k4[i], b5 = k5[i];
          ys[i] = cmake(v.x + htry * (DP_A61 * a.x + DP_A62 * b2.x + DP_A63 * b3.x + DP_A64 * b4.x + DP_A65 * b5.x),
                        v.y + htry * (DP_A61 * a.y + DP_A62 * b2.y + DP_A63 * b3.y + DP_A64 * b4.y + DP_A65 * b5.y));
        }
        __syncthreads();
        rhs<KET>(c, t + htry, ys, k6);
        for (int i = tid; i < NN; i += nth) {
          cplx v = y[i], a = k1[i], b3 = k3[i], b4 = k4[i], b5 = k5[i], b6 = k6[i];
          ys[i] = cmake(v.x + htry * (DP_B1 * a.x + DP_B3 * b3.x + DP_B4 * b4.x + DP_B5 * b5.x + DP_B6 * b6.x),
                        v.y + htry * (DP_B1 * a.y + DP_B3 * b3.y + DP_B4 * b4.y + DP_B5 * b5.y + DP_B6 * b6.y));
        }
        __syncthreads();
        rhs<KET>(c, t + htry, ys, k7);
        nrhs += 6;
        // error estimate, RMS norm scaled by atol + rtol*max(|y|,|y_new|)  (ZVODE's weighted RMS norm)
        double es = 0.0;
        for (int i = tid; i < NN; i += nth) {
          cplx a = k1[i], b3 = k3[i], b4 = k4[i], b5 = k5[i], b6 = k6[i], b7 = k7[i];
          double ex = htry * (DP_E1 * a.x + DP_E3 * b3.x + DP_E4 * b4.x + DP_E5 * b5.x + DP_E6 * b6.x + DP_E7 * b7.x);
          double ey = htry * (DP_E1 * a.y + DP_E3 * b3.y + DP_E4 * b4.y + DP_E5 * b5.y + DP_E6 * b6.y + DP_E7 * b7.y);
          double sc = p.atol + p.rtol * sqrt(fmax(cabs2(y[i]), cabs2(ys[i])));
          es += (ex * ex + ey * ey) / (sc * sc);
        }
        es = block_sum(es, c.red);
        double err = sqrt(es / (double)NN);
        if (!(err == err) || isinf(err)) { status = SEQ_STATUS_NONFINITE; break; }
        double factor;
        if (err <= 1.0) {
          factor = (err == 0.0) ? SEQ_MAX_FACTOR : fmin(SEQ_MAX_FACTOR, SEQ_SAFETY * pow(err, -0.2));
          if (prev_rejected) factor = fmin(1.0, factor);
          prev_rejected = false;
          cplx* tmp = y; y = ys; ys = tmp;
          tmp = k1; k1 = k7; k7 = tmp;
          t = landing ? t_target : t + htry;
          nacc++;
        } else {
          factor = fmax(SEQ_MIN_FACTOR, SEQ_SAFETY * pow(err, -0.2));
          prev_rejected = true;
          nrej++;
        }
        h = htry * factor;
        if (++nst > p.nsteps) { status = SEQ_STATUS_MAX_STEPS; break; }
        if (h < 1e-14 * fmax(1.0, fabs(t))) { status = SEQ_STATUS_STEP_UNDERFLOW; break; }
      }
      if (status != SEQ_STATUS_OK) break;
    }
    if (KET && (p.flags & SEQ_FLAG_NORMALIZE)) have_k1 = false;  // state is rescaled below
    emit_output<KET>(c, o, y);
    __syncthreads();
  }
  cplx* fin = p.final_state + (long long)c.b * NN;
  for (int i = tid; i < NN; i += nth) fin[i] = y[i];
  if (tid == 0) {
    p.status[c.b] = status;
    p.stats[c.b * 4 + 0] = nacc;
    p.stats[c.b * 4 + 1] = nrej;
    p.stats[c.b * 4 + 2] = nrhs;
    p.stats[c.b * 4 + 3] = 0;
  }
}

}  // namespace seq
