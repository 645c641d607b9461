// Shared device helpers: complex arithmetic on double2, the Dormand-Prince 5(4) tableau, the
// natural-cubic-spline evaluation QuTiP 4.x applies to array coefficients, block reductions,
// and the kernel parameter block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/seq_engine.h"

namespace seq {

typedef double2 cplx;

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { cplx z; z.x = r; z.y = i; return z; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return cmake(s * a.x, s * a.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// acc += a*b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc += a*conj(b)
__device__ __forceinline__ void cfma_conj(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y); acc.y = fma(-a.x, b.y, acc.y);
}
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

// ---- Dormand-Prince 5(4), FSAL.  c: nodes, a: stage matrix, b: 5th-order weights (== a7j),
// e: b - b_hat (error weights).
#define DP_C2 (1.0 / 5.0)
#define DP_C3 (3.0 / 10.0)
#define DP_C4 (4.0 / 5.0)
#define DP_C5 (8.0 / 9.0)
#define DP_A21 (1.0 / 5.0)
#define DP_A31 (3.0 / 40.0)
#define DP_A32 (9.0 / 40.0)
#define DP_A41 (44.0 / 45.0)
#define DP_A42 (-56.0 / 15.0)
#define DP_A43 (32.0 / 9.0)
#define DP_A51 (19372.0 / 6561.0)
#define DP_A52 (-25360.0 / 2187.0)
#define DP_A53 (64448.0 / 6561.0)
#define DP_A54 (-212.0 / 729.0)
#define DP_A61 (9017.0 / 3168.0)
#define DP_A62 (-355.0 / 33.0)
#define DP_A63 (46732.0 / 5247.0)
#define DP_A64 (49.0 / 176.0)
#define DP_A65 (-5103.0 / 18656.0)
#define DP_B1 (35.0 / 384.0)
#define DP_B3 (500.0 / 1113.0)
#define DP_B4 (125.0 / 192.0)
#define DP_B5 (-2187.0 / 6784.0)
#define DP_B6 (11.0 / 84.0)
#define DP_E1 (71.0 / 57600.0)
#define DP_E3 (-71.0 / 16695.0)
#define DP_E4 (71.0 / 1920.0)
#define DP_E5 (-17253.0 / 339200.0)
#define DP_E6 (22.0 / 525.0)
#define DP_E7 (-1.0 / 40.0)

// ---- embedded pairs.  pair 0: Dormand-Prince 5(4), 7 stages (6 RHS/step with FSAL);
// pair 1: the 3/8-rule RK4 with a 3rd-order FSAL error estimator (Hairer/Norsett/Wanner II.4),
// 5 stages (4 RHS/step with FSAL).  Stage s: input = y + h sum_{j<s} RK_A[pair][s][j] k_{j+1},
// k_{s+1} = f(t + RK_C[pair][s] h, input); the last row is the propagated solution (its input is
// y_new, so its derivative is k1 of the next step for either pair); error = h sum_j RK_E[pair][j] k_{j+1}.
__constant__ double RK_A[2][7][6] = {
    {{0, 0, 0, 0, 0, 0},
     {DP_A21, 0, 0, 0, 0, 0},
     {DP_A31, DP_A32, 0, 0, 0, 0},
     {DP_A41, DP_A42, DP_A43, 0, 0, 0},
     {DP_A51, DP_A52, DP_A53, DP_A54, 0, 0},
     {DP_A61, DP_A62, DP_A63, DP_A64, DP_A65, 0},
     {DP_B1, 0, DP_B3, DP_B4, DP_B5, DP_B6}},
    {{0, 0, 0, 0, 0, 0},
     {1.0 / 3.0, 0, 0, 0, 0, 0},
     {-1.0 / 3.0, 1.0, 0, 0, 0, 0},
     {1.0, -1.0, 1.0, 0, 0, 0},
     {1.0 / 8.0, 3.0 / 8.0, 3.0 / 8.0, 1.0 / 8.0, 0, 0},
     {0, 0, 0, 0, 0, 0},
     {0, 0, 0, 0, 0, 0}}};
__constant__ double RK_C[2][7] = {{0.0, DP_C2, DP_C3, DP_C4, DP_C5, 1.0, 1.0}, {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0, 1.0, 0, 0}};
__constant__ double RK_E[2][7] = {{DP_E1, 0.0, DP_E3, DP_E4, DP_E5, DP_E6, DP_E7},
                                  {1.0 / 24.0, -1.0 / 8.0, 1.0 / 8.0, 1.0 / 8.0, -1.0 / 6.0, 0, 0}};
__constant__ int RK_STAGES[2] = {7, 5};
__constant__ double RK_EXPO[2] = {-0.2, -0.25};

// The same tableaus as compile-time functions: kernels that unroll the stage loop statically get the
// coefficients as immediates and drop the zero entries.
__host__ __device__ constexpr double rk_a_c(int pair, int s, int j) {
  if (pair == 0) {
    switch (s * 8 + j) {
      case 1 * 8 + 0: return DP_A21;
      case 2 * 8 + 0: return DP_A31; case 2 * 8 + 1: return DP_A32;
      case 3 * 8 + 0: return DP_A41; case 3 * 8 + 1: return DP_A42; case 3 * 8 + 2: return DP_A43;
      case 4 * 8 + 0: return DP_A51; case 4 * 8 + 1: return DP_A52; case 4 * 8 + 2: return DP_A53; case 4 * 8 + 3: return DP_A54;
      case 5 * 8 + 0: return DP_A61; case 5 * 8 + 1: return DP_A62; case 5 * 8 + 2: return DP_A63; case 5 * 8 + 3: return DP_A64;
      case 5 * 8 + 4: return DP_A65;
      case 6 * 8 + 0: return DP_B1; case 6 * 8 + 2: return DP_B3; case 6 * 8 + 3: return DP_B4; case 6 * 8 + 4: return DP_B5;
      case 6 * 8 + 5: return DP_B6;
      default: return 0.0;
    }
  }
  switch (s * 8 + j) {
    case 1 * 8 + 0: return 1.0 / 3.0;
    case 2 * 8 + 0: return -1.0 / 3.0; case 2 * 8 + 1: return 1.0;
    case 3 * 8 + 0: return 1.0; case 3 * 8 + 1: return -1.0; case 3 * 8 + 2: return 1.0;
    case 4 * 8 + 0: return 1.0 / 8.0; case 4 * 8 + 1: return 3.0 / 8.0; case 4 * 8 + 2: return 3.0 / 8.0; case 4 * 8 + 3: return 1.0 / 8.0;
    default: return 0.0;
  }
}
__host__ __device__ constexpr double rk_e_c(int pair, int j) {
  if (pair == 0) {
    switch (j) {
      case 0: return DP_E1; case 2: return DP_E3; case 3: return DP_E4; case 4: return DP_E5; case 5: return DP_E6; case 6: return DP_E7;
      default: return 0.0;
    }
  }
  switch (j) {
    case 0: return 1.0 / 24.0; case 1: return -1.0 / 8.0; case 2: return 1.0 / 8.0; case 3: return 1.0 / 8.0; case 4: return -1.0 / 6.0;
    default: return 0.0;
  }
}
template <int V> struct IntTag { static constexpr int value = V; };

// step-size controller (same constants as scipy's RK45 / Hairer's DOPRI5)
#define SEQ_SAFETY 0.9
#define SEQ_MIN_FACTOR 0.2
#define SEQ_MAX_FACTOR 10.0
// order switching: probe the 4(3) pair when a 5(4) step was accepted at a capped step size with an
// error estimate below SEQ_PROBE_ERR; keep it only while it holds the capped step with its own
// error estimate below SEQ_STAY_ERR (otherwise 6 RHS at the capped step beat 4 RHS at a shorter one).
// Defaults of SolveParams::probe_err / stay_err (SEQ_PROBE_ERR / SEQ_STAY_ERR in the environment override them when an
// engine is created - experiments only).
#define SEQ_PROBE_ERR_DEFAULT 0.02
#define SEQ_KET_TOL_MARGIN 0.1      // internal tolerance factor of the ket path (seq_engine.cu::solve_device)
#define SEQ_STAY_ERR_DEFAULT 0.5

// ---- kernel parameter block (device pointers into caller buffers + engine scratch)
struct SolveParams {
  int N, B, n_t, n_g, n_ch, n_c, n_ctd, n_e, n_out, nsteps;
  unsigned flags;
  double t0, dt, atol, rtol, max_step, first_step, min_step;
  double probe_err, stay_err;   // order switching thresholds (see ctl_update)
  const cplx* G0;           // [N,N] (or per trajectory): H0 - (i/2) sum_const L^dag L   (kets: H0)
  long long G0_bstride;
  const cplx* Gk;           // [n_g,N,N]: H_k (k < n_ch), then -(i/2) L^dag L of time-dependent collapse ops
  const double2* tab;       // packed spline tables (y_i, M_i dt^2/6): [B or 1][n_g][n_t]
  long long tab_bstride_c;  // elements between trajectories for the coeff part (0 = shared)
  long long tab_bstride_r;  // same for the rate part
  const double2* tab_rate;  // base of the rate tables ([B or 1][n_ctd][n_t]) inside the packed buffer
  const double* coeff_scale;  // optional [B,n_ch]
  const cplx* L;            // [n_c + n_ctd, N, N]
  const cplx* state0;
  const cplx* E;
  const int* out_idx;
  void* expect;
  cplx* final_state;
  cplx* states;
  int* status;
  int* stats;
  cplx* ws;                 // per-trajectory scratch
  long long ws_stride;      // complex elements per trajectory
};

// Natural cubic spline on a uniform grid, packed table entries (y_i, m_i) with m_i = M_i dt^2/6.
// Restates qutip/cy/inter.pyx::_spline_float_cte_second: clamp outside the grid,
// value = te*(m_p te^2 + y_p - m_p) + tb*(m_{p+1} tb^2 + y_{p+1} - m_{p+1}).
__device__ __forceinline__ double spline_eval(const double2* __restrict__ tab, int n_t, double t0, double dt, double t) {
  double x = (t - t0) / dt;
  if (x <= 0.0) return tab[0].x;
  if (x >= (double)(n_t - 1)) return tab[n_t - 1].x;
  int p = (int)x;
  double tb = x - (double)p;
  double te = 1.0 - tb;
  double2 a = tab[p];
  double2 b = tab[p + 1];
  return te * (a.y * te * te + (a.x - a.y)) + tb * (b.y * tb * tb + (b.x - b.y));
}

// The same with a precomputed 1/dt (thread-per-trajectory kernel: 12 evaluations per step and lane, an FP64 division
// each).  x can differ from (t - t0) / dt by an ulp; at a knot that at most selects the neighbouring interval at its
// end point, where the C2 spline has the same value.
__device__ __forceinline__ double spline_eval_inv(const double2* __restrict__ tab, int n_t, double t0, double inv_dt, double t) {
  double x = (t - t0) * inv_dt;
  if (x <= 0.0) return tab[0].x;
  if (x >= (double)(n_t - 1)) return tab[n_t - 1].x;
  int p = (int)x;
  double tb = x - (double)p;
  double te = 1.0 - tb;
  double2 a = tab[p];
  double2 b = tab[p + 1];
  return te * (a.y * te * te + (a.x - a.y)) + tb * (b.y * tb * tb + (b.x - b.y));
}

// ---- block-wide sum; `scratch` holds >= 33 doubles.  Result is returned to every thread.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = (lane < nw) ? scratch[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) scratch[32] = s;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ cplx block_sum_c(cplx v, double* scratch) {
  double r = block_sum(v.x, scratch);
  double i = block_sum(v.y, scratch);
  return cmake(r, i);
}

// ---- per-trajectory step controller, shared by every integration kernel so that they take
// identical decisions.  All members are CTA-uniform.
struct StepCtl {
  double t, h;
  int pair;            // 0: DP5(4), 1: RK4(3)
  int status, nacc, nrej, nrhs;
  int probe_wait, probe_backoff;
  bool have_k1, prev_rejected;
};

// host copies of the tableau for the host-driven large-N stepper (same numbers as RK_A/RK_C/RK_E)
static const double H_RK_A[2][7][6] = {
    {{0, 0, 0, 0, 0, 0},
     {DP_A21, 0, 0, 0, 0, 0},
     {DP_A31, DP_A32, 0, 0, 0, 0},
     {DP_A41, DP_A42, DP_A43, 0, 0, 0},
     {DP_A51, DP_A52, DP_A53, DP_A54, 0, 0},
     {DP_A61, DP_A62, DP_A63, DP_A64, DP_A65, 0},
     {DP_B1, 0, DP_B3, DP_B4, DP_B5, DP_B6}},
    {{0, 0, 0, 0, 0, 0},
     {1.0 / 3.0, 0, 0, 0, 0, 0},
     {-1.0 / 3.0, 1.0, 0, 0, 0, 0},
     {1.0, -1.0, 1.0, 0, 0, 0},
     {1.0 / 8.0, 3.0 / 8.0, 3.0 / 8.0, 1.0 / 8.0, 0, 0},
     {0, 0, 0, 0, 0, 0},
     {0, 0, 0, 0, 0, 0}}};
static const double H_RK_C[2][7] = {{0.0, DP_C2, DP_C3, DP_C4, DP_C5, 1.0, 1.0}, {0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0, 1.0, 0, 0}};
static const double H_RK_E[2][7] = {{DP_E1, 0.0, DP_E3, DP_E4, DP_E5, DP_E6, DP_E7},
                                    {1.0 / 24.0, -1.0 / 8.0, 1.0 / 8.0, 1.0 / 8.0, -1.0 / 6.0, 0, 0}};
static const int H_RK_STAGES[2] = {7, 5};

__host__ __device__ __forceinline__ void ctl_init_at(StepCtl& c, const SolveParams& p, double t_start) {
  c.t = t_start;
  c.h = p.first_step > 0 ? p.first_step : p.dt;
  if (p.max_step > 0 && c.h > p.max_step) c.h = p.max_step;
  c.pair = 0;
  c.status = SEQ_STATUS_OK;
  c.nacc = c.nrej = c.nrhs = 0;
  c.probe_wait = 0;
  c.probe_backoff = 1;
  c.have_k1 = false;
  c.prev_rejected = false;
}
__device__ __forceinline__ void ctl_init(StepCtl& c, const SolveParams& p) {
  ctl_init_at(c, p, p.t0 + p.dt * (double)p.out_idx[0]);
}

// Next trial step towards t_target.  Returns false when the target is reached.
// `capped`: the step is limited by max_step / the output grid, not by the error controller.
__host__ __device__ __forceinline__ bool ctl_next(const StepCtl& c, const SolveParams& p, double t_target, double& htry,
                                         bool& landing, bool& capped) {
  double rem = t_target - c.t;
  if (rem <= 1e-13 * fmax(1.0, fabs(t_target))) return false;
  htry = c.h;
  if (p.max_step > 0 && htry > p.max_step) htry = p.max_step;
  double q = ceil(rem / htry - 1e-9);
  if (q < 1.0) q = 1.0;
  htry = rem / q;
  landing = (q == 1.0);
  capped = c.h > htry * (1.0 + 1e-9) || (p.max_step > 0 && c.h >= p.max_step);
  return true;
}

// Accept/reject and choose the next step size and pair.  Returns true if the step was accepted.
// ONE_POW (device only; the thread-per-trajectory kernel): the step-size factor as ONE pow() with the exponent read
// from RK_EXPO.  Where every lane carries its own controller the two-formula version below executes both formulas
// for every lane (C1: 2.96 ms instead of 2.49 ms); it differs from it by an ulp in the proposal of a 4(3) step, which
// is only taken at capped step sizes, where the proposal does not bind.
template <bool ONE_POW = false>
__host__ __device__ __forceinline__ bool ctl_update(StepCtl& c, const SolveParams& p, double err, double htry, bool landing,
                                           bool capped, double t_target, int& nst) {
  if (!(err == err) || isinf(err)) { c.status = SEQ_STATUS_NONFINITE; return false; }
  const bool switching = !(p.flags & SEQ_FLAG_NO_ORDER_SWITCH);
  bool accepted;
  if (err <= 1.0) {
    // err^(-1/4) as 1 / sqrt(sqrt(err)): correctly rounded square roots give the SAME step size on the host
    // stepper and in every kernel, and cost a fraction of pow()
    double factor;
#ifdef __CUDA_ARCH__
    if (ONE_POW) factor = (err == 0.0) ? SEQ_MAX_FACTOR : fmin(SEQ_MAX_FACTOR, SEQ_SAFETY * pow(err, RK_EXPO[c.pair]));
    else
#endif
      factor = (err == 0.0) ? SEQ_MAX_FACTOR
                            : fmin(SEQ_MAX_FACTOR, c.pair ? SEQ_SAFETY / sqrt(sqrt(err)) : SEQ_SAFETY * pow(err, -0.2));
    if (c.prev_rejected) factor = fmin(1.0, factor);
    c.prev_rejected = false;
    c.t = landing ? t_target : c.t + htry;
    c.nacc++;
    c.h = htry * factor;
    if (switching) {
      if (c.pair == 0) {
        if (c.probe_wait > 0) c.probe_wait--;
        if (capped && err <= p.probe_err && c.probe_wait == 0) c.pair = 1;
      } else if (capped && err <= p.stay_err) {
        c.probe_backoff = 1;
      } else {
        c.pair = 0;
        c.probe_backoff = (2 * c.probe_backoff < 64) ? 2 * c.probe_backoff : 64;
        c.probe_wait = c.probe_backoff;
      }
    }
    accepted = true;
  } else {
    c.nrej++;
    if (c.pair == 1) {
      // the low-order probe failed: redo the same step with the 5(4) pair, back off the next probe
      c.pair = 0;
      c.probe_backoff = (2 * c.probe_backoff < 64) ? 2 * c.probe_backoff : 64;
      c.probe_wait = c.probe_backoff;
      c.h = htry;
    } else {
#ifdef __CUDA_ARCH__
      if (ONE_POW) c.h = htry * fmax(SEQ_MIN_FACTOR, SEQ_SAFETY * pow(err, RK_EXPO[0]));
      else
#endif
        c.h = htry * fmax(SEQ_MIN_FACTOR, SEQ_SAFETY * pow(err, -0.2));
      c.prev_rejected = true;
    }
    accepted = false;
  }
  if (++nst > p.nsteps) c.status = SEQ_STATUS_MAX_STEPS;
  else if (c.h < 1e-14 * fmax(1.0, fabs(c.t))) c.status = SEQ_STATUS_STEP_UNDERFLOW;
  return accepted;
}

}  // namespace seq
