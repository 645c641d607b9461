// C-ABI implementation of include/seq_engine.h: descriptor validation, device scratch
// management, the preparation kernels (effective-Hamiltonian assembly, natural-cubic-spline
// second derivatives) and dispatch to the integration kernels.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC \
//        -o sequencing_b200/_seq_engine.so sequencing_b200/csrc/seq_engine.cu
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstddef>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "kernel_generic.cuh"
#include "kernel_dmma.cuh"
#include "kernel_dmma_stream.cuh"
#include "kernel_small.cuh"
#include "kernel_large.cuh"
#include "kernel_sparse.cuh"
#include "kernel_diag.cuh"
#include "kernel_diagblk.cuh"
#include "kernel_ket.cuh"
#include "kernel_large_diag.cuh"
#include "kernel_large_dev.cuh"
#include "kernel_large_herm.cuh"
#include "kernel_large_hermq.cuh"

using namespace seq;

// ------------------------------------------------------------------------------------------
// engine object
// ------------------------------------------------------------------------------------------
struct DevBuf {
  void* ptr = nullptr;
  size_t cap = 0;
};

struct seq_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  DevBuf ws, gbuf, tab, cprime, pack, stage_in, stage_out, large, sched, spinfo, dflags, herm;
  const seq_comm* comm = nullptr;   // set for the duration of seq_engine_solve_sharded
  void* nccl_comm = nullptr;        // native communicator (seq_engine_nccl_init), used when solve_sharded gets comm == NULL
  int nccl_rank = 0, nccl_world = 1;
  seq_comm nccl_hooks;              // hooks that issue NCCL calls on this engine's stream
  double* host_err = nullptr;       // pinned: error norm read back once per step by the large-N stepper
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int launches = 0;
  int last_method = 0;
  double probe_err = SEQ_PROBE_ERR_DEFAULT, stay_err = SEQ_STAY_ERR_DEFAULT;   // order-switch thresholds handed to every kernel
  long long last_fma_per_rhs = 0, last_fp64_step_extra = 0;   // FP64 multiply-adds per right-hand side of the last solve's kernel (0: not counted)
  std::string variant;              // human-readable detail of what the last solve ran
  int sm_count = 0;
  int max_smem_optin = 0;
  // host-buffer solves of a large batch run as chunks on two internal engines (own streams, own scratch) so that
  // the H2D / D2H copies of one chunk overlap the kernels of the other
  seq_engine* sub[2] = {nullptr, nullptr};
  cudaStream_t sub_stream[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;
  cudaStream_t aux = nullptr;       // second stream: the cluster launch of a partial last wave runs next to the main launch
  cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
  bool is_sub = false;              // an internal engine of a chunked host solve (its launches overlap the other one's)
  int chunks = 1;                   // chunks of the last host-buffer solve
  float chunk_ms = 0.f;             // their integration kernels' durations, summed
};

static int fail(seq_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf;
  return code;
}

#define CUDA_TRY(eng, call)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(eng, SEQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static int reserve(seq_engine* e, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return SEQ_OK;
  if (b.ptr) {
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaFree(b.ptr));
    b.ptr = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 8 + 256;
  cudaError_t rc = cudaMalloc(&b.ptr, want);
  if (rc != cudaSuccess) {
    b.ptr = nullptr;
    return fail(e, SEQ_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(rc));
  }
  b.cap = want;
  return SEQ_OK;
}

// ------------------------------------------------------------------------------------------
// preparation kernels
// ------------------------------------------------------------------------------------------
// K = sum_{l in [l0,l1)} L_l^dag L_l  (naive; operators are built once per solve)
__device__ __forceinline__ cplx ldagl_elem(const cplx* __restrict__ L, int N, int a, int b) {
  cplx acc = cmake(0, 0);
  for (int k = 0; k < N; k++) {
    cplx x = L[k * N + a], y = L[k * N + b];
    // conj(x) * y
    acc.x = fma(x.x, y.x, acc.x); acc.x = fma(x.y, y.y, acc.x);
    acc.y = fma(x.x, y.y, acc.y); acc.y = fma(-x.y, y.x, acc.y);
  }
  return acc;
}

// G0[b] = H0[b] - (i/2) sum_const L^dag L ; grid.y = number of G0 matrices (1 or B)
__global__ void prep_G0_kernel(const cplx* __restrict__ H0, long long h0_stride, const cplx* __restrict__ Lc, int n_c,
                               int N, cplx* __restrict__ G0, int ket) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (idx >= N * N) return;
  cplx g = H0[(long long)m * h0_stride + idx];
  if (!ket) {
    int a = idx / N, b = idx - a * N;
    cplx s = cmake(0, 0);
    for (int l = 0; l < n_c; l++) {
      cplx v = ldagl_elem(Lc + (long long)l * N * N, N, a, b);
      s.x += v.x; s.y += v.y;
    }
    // -(i/2) s = (s.y/2, -s.x/2)
    g.x += 0.5 * s.y;
    g.y -= 0.5 * s.x;
  }
  G0[(long long)m * N * N + idx] = g;
}

// Gk[m] = Hk[m] (m < n_ch) ; Gk[n_ch+j] = -(i/2) Ltd_j^dag Ltd_j
__global__ void prep_Gk_kernel(const cplx* __restrict__ Hk, int n_ch, const cplx* __restrict__ Ltd, int n_ctd, int N,
                               cplx* __restrict__ Gk) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (idx >= N * N) return;
  cplx g;
  if (m < n_ch) {
    g = Hk[(long long)m * N * N + idx];
  } else {
    int a = idx / N, b = idx - a * N;
    cplx v = ldagl_elem(Ltd + (long long)(m - n_ch) * N * N, N, a, b);
    g = cmake(0.5 * v.y, -0.5 * v.x);
  }
  Gk[(long long)m * N * N + idx] = g;
}

// Large N: row a of  out = base - (i/2) sum_{l in [0,n_ops)} L_l^dag L_l  (base == nullptr: zero), one CTA per
// (row a, output matrix).  (L^dag L)_ab = sum_k conj(L_ka) L_kb only involves the rows k in which column a is
// non-zero: the column is scanned once, its non-zero rows are listed in ascending order (deterministic sums) and
// each such row is streamed once.  O(N^2 (1 + nnz per column)) per operator instead of the O(N^3) of
// ldagl_elem - 57 ms -> 0.3 ms of set-up at C5 (N = 1875, ladder operators).
// per_op != 0: output matrix m uses operator m only (time-dependent collapse terms); else all n_ops are summed.
__global__ void prep_ldagl_rows_kernel(const cplx* __restrict__ base, const cplx* __restrict__ L, int n_ops, int per_op, int N,
                                       cplx* __restrict__ out) {
  extern __shared__ int sm_rows[];          // [N] flags, then [N] list
  int* flag = sm_rows;
  int* list = sm_rows + N;
  __shared__ int s_cnt;
  const int a = blockIdx.x, m = blockIdx.y, tid = threadIdx.x, nth = blockDim.x;
  const long long N2 = (long long)N * N;
  cplx* orow = out + (long long)m * N2 + (long long)a * N;
  for (int b = tid; b < N; b += nth) orow[b] = base ? base[(long long)m * 0 + (long long)a * N + b] : cmake(0, 0);
  const int l0 = per_op ? m : 0, l1 = per_op ? m + 1 : n_ops;
  for (int l = l0; l < l1; l++) {
    const cplx* Ll = L + (long long)l * N2;
    __syncthreads();
    for (int k = tid; k < N; k += nth) { const cplx v = Ll[(long long)k * N + a]; flag[k] = (v.x != 0.0 || v.y != 0.0) ? 1 : 0; }
    __syncthreads();
    if (tid < 32) {
      int cnt = 0;
      for (int k0 = 0; k0 < N; k0 += 32) {
        const int k = k0 + tid;
        const bool f = k < N && flag[k];
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (f) list[cnt + __popc(bal & ((1u << tid) - 1u))] = k;
        cnt += __popc(bal);
      }
      if (tid == 0) s_cnt = cnt;
    }
    __syncthreads();
    const int cnt = s_cnt;
    for (int q = 0; q < cnt; q++) {
      const int k = list[q];
      const cplx v = Ll[(long long)k * N + a];
      const cplx w = cmake(-0.5 * v.y, -0.5 * v.x);      // -(i/2) conj(v) = (-v.y/2, -v.x/2)
      for (int b = tid; b < N; b += nth) {
        const cplx x = Ll[(long long)k * N + b];
        if (x.x != 0.0 || x.y != 0.0) { cplx o = orow[b]; cfma(o, w, x); orow[b] = o; }
      }
    }
  }
}

// Thomas factors of the (1,4,1) system, shared by every series: cp[i] for i in 1..n_t-2
__global__ void spline_cprime_kernel(double* cp, int n_t) {
  if (blockIdx.x || threadIdx.x) return;
  double c = 0.0;
  for (int i = 1; i <= n_t - 2; i++) {
    c = 1.0 / (4.0 - c);
    cp[i] = c;
  }
}

// One thread per series: m solves m[i-1] + 4 m[i] + m[i+1] = y[i+1] - 2 y[i] + y[i-1], m[0]=m[n-1]=0
// (m = M dt^2/6 of QuTiP's _prep_cubic_spline).  Output is packed (y, m) per sample.
__global__ void spline_series_kernel(const double* __restrict__ y, double2* __restrict__ out, const double* __restrict__ cp,
                                     long long n_series, int n_t) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_series) return;
  const double* ys = y + s * n_t;
  double2* o = out + s * n_t;
  if (n_t < 3) {
    for (int i = 0; i < n_t; i++) o[i] = make_double2(ys[i], 0.0);
    return;
  }
  double dprev = 0.0;
  double ym = ys[0], yc = ys[1];
  o[0] = make_double2(ym, 0.0);
  for (int i = 1; i <= n_t - 2; i++) {
    double yp = ys[i + 1];
    double rhs = yp - 2.0 * yc + ym;
    double d = (rhs - dprev) * cp[i];
    o[i] = make_double2(yc, d);
    dprev = d;
    ym = yc;
    yc = yp;
  }
  o[n_t - 1] = make_double2(yc, 0.0);
  double mnext = 0.0;
  for (int i = n_t - 2; i >= 1; i--) {
    double2 v = o[i];
    double m = v.y - cp[i] * mnext;
    o[i] = make_double2(v.x, m);
    mnext = m;
  }
}

// unpack helper for seq_engine_spline_prep: M[i] = m[i] * 6 / dt^2
__global__ void spline_unpack_kernel(const double2* __restrict__ packed, double* __restrict__ M, long long n, double scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) M[i] = packed[i].y * scale;
}

// One CTA per pulse descriptor: chopped Gaussian(-DRAG) waveform, modulated and added to the coefficient tables
// (include/seq_engine.h: seq_engine_synth_gaussian).  linspace() is evaluated as numpy does (start + k step, last
// sample exactly the stop value) so that the samples agree with pulses.py to the last bits of exp / sincos.
__global__ void synth_gaussian_kernel(const seq_pulse_desc* __restrict__ descs, double* __restrict__ coeff, int B, int n_ch, int n_t) {
  const seq_pulse_desc d = descs[blockIdx.x];
  const double len = d.chop * d.sigma;
  const int n = (int)(len / d.dt);
  if (d.traj < 0 || d.traj >= B || n < 1) return;
  const double t_lo = -len / 2, t_hi = len / 2;
  const double t_step = n > 1 ? (t_hi - t_lo) / (double)(n - 1) : 0.0;
  const double tau_hi = (double)n * d.dt, tau_step = n > 1 ? tau_hi / (double)(n - 1) : 0.0;
  const double inv2s2 = 1.0 / (2.0 * d.sigma * d.sigma);
  const double g0 = exp(-(t_lo * t_lo) * inv2s2);
  const double cph = cos(d.phase), sph = sin(d.phase);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int col = d.start + k;
    if (col < 0 || col >= n_t) continue;
    const double t = (k == n - 1 && n > 1) ? t_hi : t_lo + (double)k * t_step;
    const double g = exp(-(t * t) * inv2s2);
    double wr = (g - g0) / (1.0 - g0);
    double wi = d.drag * ((0.25 / (d.sigma * d.sigma)) * t * g / (1.0 - g0));
    if (d.detune != 0.0) {
      const double tau = (k == n - 1 && n > 1) ? tau_hi : (double)k * tau_step;
      const double ang = -2.0 * 3.141592653589793238462643383279502884 * tau * d.detune;
      const double c = cos(ang), s_ = sin(ang);
      const double r = wr * c - wi * s_, i = wr * s_ + wi * c;
      wr = r; wi = i;
    }
    wr *= d.amp; wi *= d.amp;
    const double re = wr * cph - wi * sph, im = wr * sph + wi * cph;
    if (d.ch_re >= 0 && d.ch_re < n_ch) atomicAdd(&coeff[((long long)d.traj * n_ch + d.ch_re) * n_t + col], d.sign_re * re);
    if (d.ch_im >= 0 && d.ch_im < n_ch) atomicAdd(&coeff[((long long)d.traj * n_ch + d.ch_im) * n_t + col], d.sign_im * im);
  }
}

// states[b] <- U states[b] U^dag  /  U psi_b   (one CTA per trajectory; scratch T in global)
__global__ void apply_unitary_kernel(const cplx* __restrict__ U, cplx* __restrict__ states, cplx* __restrict__ scratch,
                                     int N, int is_ket) {
  int b = blockIdx.x;
  if (is_ket) {
    cplx* psi = states + (long long)b * N;
    cplx* T = scratch + (long long)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      cplx acc = cmake(0, 0);
      for (int k = 0; k < N; k++) cfma(acc, U[i * N + k], psi[k]);
      T[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) psi[i] = T[i];
    return;
  }
  cplx* rho = states + (long long)b * N * N;
  cplx* T = scratch + (long long)b * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    int i = idx / N, j = idx - i * N;
    cplx acc = cmake(0, 0);
    for (int k = 0; k < N; k++) cfma(acc, U[i * N + k], rho[k * N + j]);
    T[idx] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    int i = idx / N, j = idx - i * N;
    cplx acc = cmake(0, 0);
    for (int k = 0; k < N; k++) cfma_conj(acc, T[i * N + k], U[j * N + k]);
    rho[idx] = acc;
  }
}

__global__ void expect_kernel(const cplx* __restrict__ E, const cplx* __restrict__ states, cplx* __restrict__ out, int N,
                              int n_e, int is_ket) {
  __shared__ double red[34];
  int b = blockIdx.x, e = blockIdx.y;
  const cplx* Em = E + (long long)e * N * N;
  cplx acc = cmake(0, 0);
  if (is_ket) {
    const cplx* psi = states + (long long)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      cplx row = cmake(0, 0);
      for (int k = 0; k < N; k++) cfma(row, Em[i * N + k], psi[k]);
      cfma_conj(acc, row, psi[i]);
    }
  } else {
    const cplx* rho = states + (long long)b * N * N;
    for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
      int i = idx / N, j = idx - i * N;
      cfma(acc, Em[idx], rho[j * N + i]);
    }
  }
  acc = block_sum_c(acc, red);
  if (threadIdx.x == 0) out[(long long)b * n_e + e] = acc;
}

// ------------------------------------------------------------------------------------------
// descriptor checks and method selection
// ------------------------------------------------------------------------------------------
static int validate(seq_engine* e, const seq_solve_desc* d) {
  if (!d) return fail(e, SEQ_ERR_INVALID, "descriptor is NULL");
  if (d->size != sizeof(seq_solve_desc))
    return fail(e, SEQ_ERR_INVALID, "descriptor size %u != %zu (ABI mismatch)", d->size, sizeof(seq_solve_desc));
  if (d->N < 1 || d->B < 0 || d->n_t < 1 || d->n_out < 1)
    return fail(e, SEQ_ERR_INVALID, "bad shape N=%d B=%d n_t=%d n_out=%d", d->N, d->B, d->n_t, d->n_out);
  if (d->n_ch < 0 || d->n_c < 0 || d->n_ctd < 0 || d->n_e < 0)
    return fail(e, SEQ_ERR_INVALID, "negative operator count");
  if (!(d->dt > 0)) return fail(e, SEQ_ERR_INVALID, "dt must be positive");
  if (!(d->atol > 0) || !(d->rtol >= 0)) return fail(e, SEQ_ERR_INVALID, "bad tolerances atol=%g rtol=%g", d->atol, d->rtol);
  if (d->nsteps < 1) return fail(e, SEQ_ERR_INVALID, "nsteps must be >= 1");
  if (d->B == 0) return SEQ_OK;
  if (!d->H0 || !d->state0 || !d->final_state || !d->status || !d->stats || !d->out_idx)
    return fail(e, SEQ_ERR_INVALID, "required pointer is NULL (H0/state0/final_state/status/stats/out_idx)");
  if (d->n_ch > 0 && (!d->Hk || !d->coeff)) return fail(e, SEQ_ERR_INVALID, "n_ch > 0 but Hk/coeff is NULL");
  if (d->n_c > 0 && !d->Lc) return fail(e, SEQ_ERR_INVALID, "n_c > 0 but Lc is NULL");
  if (d->n_ctd > 0 && (!d->Ltd || !d->rate)) return fail(e, SEQ_ERR_INVALID, "n_ctd > 0 but Ltd/rate is NULL");
  if (d->n_e > 0 && (!d->E || !d->expect)) return fail(e, SEQ_ERR_INVALID, "n_e > 0 but E/expect is NULL");
  if ((d->flags & SEQ_FLAG_STORE_STATES) && !d->states)
    return fail(e, SEQ_ERR_INVALID, "SEQ_FLAG_STORE_STATES set but states is NULL");
  if ((d->flags & SEQ_FLAG_KET) && (d->n_c + d->n_ctd) > 0)
    return fail(e, SEQ_ERR_INVALID, "kets cannot be evolved with collapse operators (convert to a density matrix)");
  if (d->n_ch + d->n_ctd > 64) return fail(e, SEQ_ERR_UNSUPPORTED, "more than 64 time-dependent terms");
  return SEQ_OK;
}

// out_idx: strictly increasing sample indices inside [0, n_t) - the kernels index the spline tables and the output
// arrays with them.  Checked on the host copy (host-pointer solves) or after one small D2H copy (device pointers).
static int validate_out_idx(seq_engine* e, const int32_t* idx, int n_out, int n_t, bool has_tables) {
  for (int o = 0; o < n_out; o++) {
    if (idx[o] < 0 || (has_tables && idx[o] >= n_t))
      return fail(e, SEQ_ERR_INVALID, "out_idx[%d] = %d is outside the coefficient grid [0, %d)", o, idx[o], n_t);
    if (o > 0 && idx[o] <= idx[o - 1])
      return fail(e, SEQ_ERR_INVALID, "out_idx must be strictly increasing (out_idx[%d] = %d after %d)", o, idx[o], idx[o - 1]);
  }
  return SEQ_OK;
}

static int pick_method(const seq_engine* e, const seq_solve_desc* d) {
  if (d->method != SEQ_METHOD_AUTO) return d->method;
  const bool ket = d->flags & SEQ_FLAG_KET;
  if (!ket && d->H0_batch_stride == 0 && small_supported(d->N, d->n_ch + d->n_ctd, d->n_e)) return SEQ_METHOD_SMALL;
  const int n_l = d->n_c + d->n_ctd;
  if (!ket && d->H0_batch_stride == 0 && n_l <= DmmaCfg<6>::MAXL && dmma_supported(d->N, e ? e->max_smem_optin : 232448)) return SEQ_METHOD_DMMA;
  if (!ket && d->H0_batch_stride == 0 && n_l <= StreamCfg<6>::MAXL && stream_supported(d->N, e ? e->max_smem_optin : 232448)) return SEQ_METHOD_DMMA_STREAM;
  if (!ket && d->H0_batch_stride == 0 && d->N > 96) return SEQ_METHOD_LARGE;
  return SEQ_METHOD_GENERIC;
}

static size_t state_elems(const seq_solve_desc* d) {
  return (d->flags & SEQ_FLAG_KET) ? (size_t)d->N : (size_t)d->N * d->N;
}

static size_t ws_elems_per_traj(const seq_solve_desc* d, int method) {
  size_t N2 = (size_t)d->N * d->N, NN = state_elems(d);
  if (method == SEQ_METHOD_GENERIC) return 2 * N2 + 9 * NN;
  if (method == SEQ_METHOD_DMMA) return dmma_ws_elems(d->N);
  if (method == SEQ_METHOD_DMMA_STREAM) return stream_ws_elems(d->N);
  return 0;
}

// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen of the library the host process already carries): no link-time dependency
// ------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };      // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool nccl_load() {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!h) continue;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.CommDestroy) { g_nccl.lib = h; return true; }
    dlclose(h);
  }
  return false;
}
#define SEQ_NCCL_DOUBLE 8   /* ncclFloat64 */
#define SEQ_NCCL_SUM 0      /* ncclSum */
static int nccl_hook_all_gather(void* ctx, const double* send, double* recv, int64_t count) {
  seq_engine* e = (seq_engine*)ctx;
  return g_nccl.AllGather(send, recv, (size_t)count, SEQ_NCCL_DOUBLE, e->nccl_comm, e->stream);
}
static int nccl_hook_all_reduce(void* ctx, double* buf, int64_t count) {
  seq_engine* e = (seq_engine*)ctx;
  return g_nccl.AllReduce(buf, buf, (size_t)count, SEQ_NCCL_DOUBLE, SEQ_NCCL_SUM, e->nccl_comm, e->stream);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int seq_engine_nccl_unique_id(void* id_out) {
  if (!id_out || !nccl_load()) return SEQ_ERR_UNSUPPORTED;
  return g_nccl.GetUniqueId(id_out) == 0 ? SEQ_OK : SEQ_ERR_CUDA;
}

int seq_engine_nccl_init(seq_engine* e, const void* id, int32_t rank, int32_t world) {
  if (!e || !id || world < 1 || rank < 0 || rank >= world) return SEQ_ERR_INVALID;
  if (!nccl_load()) return fail(e, SEQ_ERR_UNSUPPORTED, "libnccl.so.2 could not be opened: %s", dlerror());
  if (e->nccl_comm) { g_nccl.CommDestroy(e->nccl_comm); e->nccl_comm = nullptr; }
  CUDA_TRY(e, cudaSetDevice(e->device));
  NcclId nid;
  memcpy(&nid, id, sizeof(nid));
  const int rc = g_nccl.CommInitRank(&e->nccl_comm, world, nid, rank);
  if (rc != 0) { e->nccl_comm = nullptr; return fail(e, SEQ_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); }
  e->nccl_rank = rank; e->nccl_world = world;
  e->nccl_hooks.rank = rank; e->nccl_hooks.world = world; e->nccl_hooks.ctx = e;
  e->nccl_hooks.all_gather = nccl_hook_all_gather;
  e->nccl_hooks.all_reduce_sum = nccl_hook_all_reduce;
  return SEQ_OK;
}

int seq_engine_nccl_destroy(seq_engine* e) {
  if (!e) return SEQ_ERR_INVALID;
  if (e->nccl_comm && g_nccl.lib) { cudaStreamSynchronize(e->stream); g_nccl.CommDestroy(e->nccl_comm); }
  e->nccl_comm = nullptr;
  e->nccl_world = 1;
  return SEQ_OK;
}

int seq_engine_abi_version(void) { return SEQ_ENGINE_ABI_VERSION; }

const char* seq_engine_build_info(void) {
  return "seq_engine abi=3 arch=sm_100a kernels=diagblk(operators by diagonals, one tensor factor inside a thread, fma),diag(operators by diagonals, fma, rho upper triangle in registers),ketdiag(kets by diagonals),sparse(ELL operator rows, fma),generic(fma),dmma(8x8x4 resident),dmma_stream(8x8x4 blocked),small(thread-per-trajectory real superoperator),large(Hermitian tiles with fused stage kernels, outer tensor factor inside a thread, device-resident step control | full-matrix structured RHS | stream-K DMMA ZGEMM; row-sharded over hooks or native NCCL) built " __DATE__;
}

int seq_engine_create(int device, void* stream, seq_engine** out) {
  if (!out) return SEQ_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t rc = cudaGetDeviceCount(&count);
  if (rc != cudaSuccess || device < 0 || device >= count) return SEQ_ERR_CUDA;
  seq_engine* e = new seq_engine();
  e->device = device;
  e->stream = (cudaStream_t)stream;
  if (cudaSetDevice(device) != cudaSuccess) { delete e; return SEQ_ERR_CUDA; }
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&e->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess) { delete e; return SEQ_ERR_CUDA; }
  *out = e;
  return SEQ_OK;
}

int seq_engine_destroy(seq_engine* e) {
  if (!e) return SEQ_OK;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->nccl_comm && g_nccl.lib) { g_nccl.CommDestroy(e->nccl_comm); e->nccl_comm = nullptr; }
  DevBuf* bufs[] = {&e->ws, &e->gbuf, &e->tab, &e->cprime, &e->pack, &e->stage_in, &e->stage_out, &e->large, &e->sched, &e->spinfo, &e->dflags, &e->herm};
  for (DevBuf* b : bufs) if (b->ptr) cudaFree(b->ptr);
  if (e->host_err) cudaFreeHost(e->host_err);
  for (int q = 0; q < 2; q++) {
    if (e->sub[q]) seq_engine_destroy(e->sub[q]);
    if (e->sub_stream[q]) cudaStreamDestroy(e->sub_stream[q]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->aux) cudaStreamDestroy(e->aux);
  if (e->ev_aux0) cudaEventDestroy(e->ev_aux0);
  if (e->ev_aux1) cudaEventDestroy(e->ev_aux1);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  delete e;
  return SEQ_OK;
}

const char* seq_engine_last_error(const seq_engine* e) { return e ? e->err.c_str() : "engine is NULL"; }

int seq_engine_select_method(const seq_engine* e, const seq_solve_desc* d) {
  if (!d) return SEQ_ERR_INVALID;
  return pick_method(e, d);
}

int64_t seq_engine_workspace_bytes(const seq_engine* e, const seq_solve_desc* d) {
  if (!d || d->size != sizeof(seq_solve_desc)) return SEQ_ERR_INVALID;
  int method = pick_method(e, d);
  size_t N2 = (size_t)d->N * d->N;
  int n_g = d->n_ch + d->n_ctd;
  size_t nG0 = d->H0_batch_stride ? (size_t)d->B : 1;
  size_t bytes = ws_elems_per_traj(d, method) * (size_t)d->B * sizeof(cplx);
  bytes += (nG0 + n_g) * N2 * sizeof(cplx);
  size_t series_c = (size_t)(d->coeff_batch_stride ? d->B : 1) * d->n_ch;
  size_t series_r = (size_t)(d->rate_batch_stride ? d->B : 1) * d->n_ctd;
  bytes += (series_c + series_r) * (size_t)d->n_t * sizeof(double2) + (size_t)d->n_t * sizeof(double);
  return (int64_t)bytes;
}

int seq_engine_last_launch_count(const seq_engine* e) { return e ? e->launches : 0; }

int seq_engine_last_method(const seq_engine* e) { return e ? e->last_method : 0; }

const char* seq_engine_last_variant(const seq_engine* e) { return e ? e->variant.c_str() : ""; }

int64_t seq_engine_last_fma_per_rhs(const seq_engine* e) { return e ? (int64_t)e->last_fma_per_rhs : 0; }
int64_t seq_engine_last_step_extra_ops(const seq_engine* e) { return e ? (int64_t)e->last_fp64_step_extra : 0; }

float seq_engine_last_kernel_ms(seq_engine* e) {
  if (e && e->chunks > 1) return e->chunk_ms;
  if (!e || !e->timed) return -1.0f;
  if (cudaEventSynchronize(e->ev1) != cudaSuccess) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) return -1.0f;
  return ms;
}

static int spline_prep_packed(seq_engine* e, const double* y, double2* packed, long long n_series, int n_t) {
  if (n_series == 0) return SEQ_OK;
  int rc = reserve(e, e->cprime, (size_t)n_t * sizeof(double));
  if (rc) return rc;
  spline_cprime_kernel<<<1, 32, 0, e->stream>>>((double*)e->cprime.ptr, n_t);
  e->launches++;
  int threads = 128;
  long long blocks = (n_series + threads - 1) / threads;
  spline_series_kernel<<<(unsigned)blocks, threads, 0, e->stream>>>(y, packed, (const double*)e->cprime.ptr, n_series, n_t);
  e->launches++;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

int seq_engine_spline_prep(seq_engine* e, const double* y, double* M, int64_t n_series, int32_t n_t, double dt) {
  if (!e) return SEQ_ERR_INVALID;
  if (!y || !M || n_series < 0 || n_t < 1 || !(dt > 0)) return fail(e, SEQ_ERR_INVALID, "bad spline_prep arguments");
  CUDA_TRY(e, cudaSetDevice(e->device));
  e->launches = 0;
  size_t n = (size_t)n_series * n_t;
  if (n == 0) return SEQ_OK;
  int rc = reserve(e, e->tab, n * sizeof(double2));
  if (rc) return rc;
  rc = spline_prep_packed(e, y, (double2*)e->tab.ptr, n_series, n_t);
  if (rc) return rc;
  spline_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>((const double2*)e->tab.ptr, M, (long long)n, 6.0 / (dt * dt));
  e->launches++;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

// ------------------------------------------------------------------------------------------
// SEQ_METHOD_LARGE: host-driven embedded-RK stepper over grid-wide kernels (kernel_large.cuh).
// One trajectory at a time; with a communicator, rho's rows are sharded over the ranks and the
// stage state is all-gathered once per RHS evaluation.
// ------------------------------------------------------------------------------------------
static int grid_for(long long n, int sms) {
  long long g = (n + 255) / 256;
  long long cap = (long long)sms * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

__global__ void large_store_expect_kernel(const cplx* __restrict__ acc, void* __restrict__ expect, long long base, long long stride,
                                          int n_e, int complex_out) {
  int e = threadIdx.x;
  if (e >= n_e) return;
  if (complex_out) ((cplx*)expect)[base + e * stride] = acc[e];
  else ((double*)expect)[base + e * stride] = acc[e].x;
}

// Row statistics of the operators (device) -> ELL widths; decides whether SEQ_METHOD_SPARSE applies.
// One small D2H copy + stream synchronisation.
struct SparseInfo {
  bool usable = false;       // widths within the kernel's limits and a launch plan exists
  bool worthwhile = false;   // fewer multiply-adds than the dense tensor-core formulation by a wide margin
  SparsePlan plan;
  int* rowcnt = nullptr;     // device: [n_ops][N]
};

static int sparse_analyze(seq_engine* e, const seq_solve_desc* d, const cplx* G0, const cplx* Gk, const cplx* Lall, SparseInfo* info) {
  const int N = d->N, n_g = d->n_ch + d->n_ctd, n_l = d->n_c + d->n_ctd, n_e = d->n_e;
  const int n_ops = 1 + n_l + n_e;
  size_t ints = (size_t)n_ops * N + 2 * (size_t)n_ops + (size_t)(n_e > 0 ? n_e : 1) * N + 16;
  int rc = reserve(e, e->spinfo, ints * sizeof(int));
  if (rc) return rc;
  int* rowcnt = (int*)e->spinfo.ptr;
  int* maxw = rowcnt + (size_t)n_ops * N;
  int* total = maxw + n_ops;
  CUDA_TRY(e, cudaMemsetAsync(maxw, 0, 2 * (size_t)n_ops * sizeof(int), e->stream));
  dim3 grid(N, n_ops);
  sparse_count_kernel<<<grid, 32, 0, e->stream>>>(G0, Gk, n_g, Lall, n_l, (const cplx*)d->E, N, rowcnt, maxw, total);
  e->launches++;
  std::vector<int> h(2 * (size_t)n_ops);
  CUDA_TRY(e, cudaMemcpyAsync(h.data(), maxw, h.size() * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  int wL = 1;
  for (int l = 0; l < n_l; l++) wL = h[1 + l] > wL ? h[1 + l] : wL;
  long long nnzE = 0;
  for (int q = 0; q < n_e; q++) nnzE += h[n_ops + 1 + n_l + q];
  info->rowcnt = rowcnt;
  const int wG = h[0] < 1 ? 1 : h[0];
  info->worthwhile = sparse_worthwhile(N, n_l, wG, h.data() + 1);
  info->plan.nnzE = (int)nnzE;
  info->usable = wG <= SPARSE_MAX_WG && wL <= SPARSE_MAX_WL && nnzE < (1LL << 30) &&
                 sparse_plan(N, n_g, n_l, wG, wL, (int)nnzE, e->max_smem_optin, &info->plan);
  info->plan.nnzE = (int)nnzE;
  info->plan.macs = 2LL * wG;
  for (int l = 0; l < n_l; l++) info->plan.macs += (long long)h[1 + l] * h[1 + l];
  return SEQ_OK;
}

static int solve_large(seq_engine* e, const seq_solve_desc* d, const SolveParams& p, const cplx* G0, const cplx* Gk) {
  const seq_comm* comm = e->comm;
  const int world = (comm && comm->world > 1) ? comm->world : 1;
  const int rank = world > 1 ? comm->rank : 0;
  const int N = d->N, NP = large_np(N), B = d->B;
  const int n_g = p.n_g, n_l = p.n_c + p.n_ctd, n_e = p.n_e;
  const int m = ((N + world - 1) / world + 15) / 16 * 16;   // rows per rank
  const long long rows_full = (long long)world * m;          // >= NP
  const long long row0 = (long long)rank * m;
  const int Mrows = (int)((N - row0) < 0 ? 0 : ((N - row0) > m ? m : (N - row0)));
  const long long opsz = 2LL * NP * NP, ysz = 2LL * rows_full * NP, lsz = 2LL * m * NP;
  const long long mplane = (long long)m * NP, yplane = rows_full * NP, oplane = (long long)NP * NP;
  cudaStream_t st = e->stream;
  const int sms = e->sm_count;

  // ---- operator structure: sums of a few diagonals -> one structured RHS kernel instead of the ZGEMM chain
  bool structured = false;
  LargeDiagArgs la;
  DiagFill dfill;
  memset(&la, 0, sizeof(la));
  int rc;
  if (!getenv("SEQ_LARGE_DENSE")) {
    const size_t W = 2 * (size_t)N - 1, nflag = (size_t)(2 + n_l) * W;
    rc = reserve(e, e->dflags, nflag * sizeof(int) + 64);
    if (rc) return rc;
    int* flags = (int*)e->dflags.ptr;
    CUDA_TRY(e, cudaMemsetAsync(flags, 0, nflag * sizeof(int), st));
    dim3 fg((unsigned)grid_for((long long)N * N, sms), 2 + n_l);
    diag_flags_kernel<<<fg, 256, 0, st>>>(G0, Gk, n_g, p.L, N, flags);
    e->launches++;
    std::vector<int> hflags(nflag);
    CUDA_TRY(e, cudaMemcpyAsync(hflags.data(), flags, nflag * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    structured = large_diag_terms(N, p.n_c, p.n_ctd, p.n_ch, hflags.data(), &la, &dfill);
  }
  e->variant = structured ? "large/diagonals" : "large/zgemm";
  cplx *dg0 = nullptr, *dgk = nullptr, *dlam = nullptr, *dgdiag = nullptr, *dgt = nullptr;
  if (structured) {
    const size_t nDN = (size_t)dfill.nD * N, nTN = (size_t)(dfill.nT > 0 ? dfill.nT : 1) * N;
    rc = reserve(e, e->pack, ((2 + (size_t)(n_g > 0 ? n_g : 1)) * nDN + nTN + N) * sizeof(cplx) + 256);
    if (rc) return rc;
    dg0 = (cplx*)e->pack.ptr; dgk = dg0 + nDN; dlam = dgk + (size_t)(n_g > 0 ? n_g : 1) * nDN; dgdiag = dlam + nTN; dgt = dgdiag + N;
    rc = reserve(e, e->spinfo, 64);
    if (rc) return rc;
    dim3 gf((N + 127) / 128, dfill.nD + dfill.nT);
    diag_fill_kernel<<<gf, 128, 0, st>>>(dfill, G0, Gk, n_g, p.L, N, dg0, dgk, dlam, dgdiag, (int*)e->spinfo.ptr);
    e->launches++;
  }

  size_t total = (size_t)(structured ? 0 : (1 + n_g + n_l + 1)) * opsz + ysz + (world > 1 ? lsz : 0) + (size_t)(2 + 7 + 1) * lsz + 64 + 2 + 2 * (size_t)(n_e + 1);
  rc = reserve(e, e->large, total * sizeof(double) + 256);
  if (rc) return rc;
  if (!e->host_err) CUDA_TRY(e, cudaMallocHost((void**)&e->host_err, sizeof(LargeDev) + 64));      // error norm / control block read-backs
  double* base = (double*)e->large.ptr;
  double *G0p = nullptr, *Gkp = nullptr, *Lp = nullptr, *Gt = nullptr;
  if (!structured) {
    G0p = base; base += opsz;
    Gkp = base; base += (size_t)n_g * opsz;
    Lp = base; base += (size_t)n_l * opsz;
    Gt = base; base += opsz;
  }
  double* Yfull = base; base += ysz;
  double* Ysend = nullptr;
  if (world > 1) { Ysend = base; base += lsz; }
  double* ycur = base; base += lsz;
  double* ynew = base; base += lsz;
  double* kbuf[7];
  for (int j = 0; j < 7; j++) { kbuf[j] = base; base += lsz; }
  double* T = base; base += lsz;
  double* cval = base; base += 64;
  double* derr = base; base += 2;
  cplx* eacc = (cplx*)base;

  // operators -> planar padded
  if (!structured) {
  large_pack_kernel<<<grid_for(oplane, sms), 256, 0, st>>>(G0, G0p, N, NP, NP);
  e->launches++;
  for (int g = 0; g < n_g; g++) { large_pack_kernel<<<grid_for(oplane, sms), 256, 0, st>>>(Gk + (size_t)g * N * N, Gkp + (size_t)g * opsz, N, NP, NP); e->launches++; }
  for (int l = 0; l < n_l; l++) { large_pack_kernel<<<grid_for(oplane, sms), 256, 0, st>>>(p.L + (size_t)l * N * N, Lp + (size_t)l * opsz, N, NP, NP); e->launches++; }
  }

  // chunk schedules: per operator (0: union pattern of G0 and every Gk; 1..: L_l) one list for its use as
  // the A operand (48-row tiles of this rank's rows) and one for its use as the B^H operand (96-row tiles)
  const int kchunks = NP / LargeGemmCfg::KC;
  const int tilesA = (Mrows + LargeGemmCfg::TM - 1) / LargeGemmCfg::TM, tilesB = (NP + LargeGemmCfg::TN - 1) / LargeGemmCfg::TN;
  const int nops = 1 + n_l;
  const size_t idx_per_op = (size_t)(tilesA + tilesB) * kchunks;          // unsigned short
  const size_t int_per_op = (size_t)2 * (tilesA + tilesB) + 4;            // cnt + pref
  rc = reserve(e, e->sched, nops * (idx_per_op * sizeof(unsigned short) + int_per_op * sizeof(int)) + 256);
  if (rc) return rc;
  int* s_int = (int*)e->sched.ptr;
  unsigned short* s_idx = (unsigned short*)(s_int + nops * int_per_op);
  auto idxA = [&](int op) { return s_idx + (size_t)op * idx_per_op; };
  auto idxB = [&](int op) { return idxA(op) + (size_t)tilesA * kchunks; };
  auto prefA = [&](int op) { return s_int + (size_t)op * int_per_op + (tilesA + tilesB); };
  auto prefB = [&](int op) { return prefA(op) + tilesA + 1; };
  for (int op = 0; op < (structured ? 0 : nops); op++) {
    const double* src = (op == 0) ? G0p : Lp + (size_t)(op - 1) * opsz;
    const int nsrc = (op == 0) ? 1 + n_g : 1;     // G0p and Gkp are contiguous
    int* cntA = s_int + (size_t)op * int_per_op;
    int* cntB = cntA + tilesA;
    if (tilesA > 0) {
      large_sched_kernel<<<tilesA, 128, kchunks, st>>>(src, opsz, nsrc, NP, oplane, row0, Mrows, LargeGemmCfg::TM, kchunks, idxA(op), cntA);
      e->launches++;
    }
    large_sched_kernel<<<tilesB, 128, kchunks, st>>>(src, opsz, nsrc, NP, oplane, 0, NP, LargeGemmCfg::TN, kchunks, idxB(op), cntB);
    large_pref_kernel<<<1, 32, 0, st>>>(cntA, tilesA, prefA(op));
    large_pref_kernel<<<1, 32, 0, st>>>(cntB, tilesB, prefB(op));
    e->launches += 3;
  }

  std::vector<int> out_idx(d->n_out);
  CUDA_TRY(e, cudaMemcpyAsync(out_idx.data(), d->out_idx, sizeof(int) * d->n_out, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(e, cudaStreamSynchronize(st));

  // ---- one GPU + structured operators: the step controller lives on the device (kernel_large_dev.cuh) - no per-step
  // D2H copy / synchronisation; the host enqueues attempts ahead of the GPU.  SEQ_LARGE_HOST=1 keeps the host-driven loop.
  const bool devctl = structured && world == 1 && n_g <= LDEV_MAXG && N <= 0xffff && !getenv("SEQ_LARGE_HOST");
  LargeDev* ddev = nullptr;
  int *e_off = nullptr; unsigned* e_ij = nullptr; cplx* e_val = nullptr;
  if (devctl) {
    SparseInfo spi;
    rc = sparse_analyze(e, d, G0, Gk, p.L, &spi);      // row counts of the expectation operators -> COO lists
    if (rc) return rc;
    SparsePlan epl;
    memset(&epl, 0, sizeof(epl));
    epl.wG = 1; epl.wL = 1; epl.nnzE = spi.plan.nnzE;
    SparsePack pk = sparse_pack_layout(N, 0, 0, n_e, epl);
    rc = reserve(e, e->sched, pk.bytes + sizeof(LargeDev) + 512);
    if (rc) return rc;
    char* sb = (char*)e->sched.ptr;
    e_off = (int*)(sb + pk.o_eoff); e_ij = (unsigned*)(sb + pk.o_eij); e_val = (cplx*)(sb + pk.o_eval);
    ddev = (LargeDev*)(sb + (pk.bytes + 255) / 256 * 256);
    const int n_ops = 1 + n_l + n_e;
    int* rowoff = spi.rowcnt + (size_t)n_ops * N + 2 * (size_t)n_ops;
    sparse_escan_kernel<<<1, 32, 0, st>>>(spi.rowcnt, n_l, n_e, N, rowoff, e_off);
    e->launches++;
    if (n_e > 0) {
      dim3 ge(N, n_e);
      sparse_fill_kernel<<<ge, 32, 0, st>>>(G0, Gk, n_g, p.L, n_l, p.E, N, 1, 1, nullptr, nullptr, nullptr, nullptr, nullptr, rowoff, e_ij, e_val, 1 + n_l);
      e->launches++;
    }
    e->variant = "large/diagonals (device-resident step control)";
  }
  // ---- Hermitian tiles (kernel_large_herm.cuh): half of rho, stage combination / error estimate fused into the RHS kernel.
  // SEQ_LARGE_FULL=1 keeps the full-matrix kernels; a non-Hermitian rho0 (checked per trajectory) takes them as well.
  const bool herm_try = devctl && !getenv("SEQ_LARGE_FULL");
  LargeHermBufs hb;
  memset(&hb, 0, sizeof(hb));
  int h_tiles = 0, NPh = 0;
  size_t h_smem = 0, hq_smem_t = 0;      // dynamic shared memory of the stage kernel / of the q-local pack and stage-input kernels
  void (*h_stage)(LargeDev*, const LargeHermBufs, int, const LargeDiagArgs, double, double) = nullptr;
  // q-local variant (kernel_large_hermq.cuh): the outermost tensor factor inside a thread
  LargeQLocal lq;
  bool use_q = false;
  const bool h_pdl = !getenv("SEQ_LH_NOPDL");
  void (*hq_stage)(LargeDev*, const LargeHermBufs, int, const LargeDiagArgs, const LargeQLocal, double, double) = nullptr;
  void (*hq_first)(const LargeDev*, const LargeHermBufs, int, int, const LargeHermExpect) = nullptr;
  void (*hq_pack)(const cplx*, const LargeHermBufs, int, unsigned long long*) = nullptr;
  unsigned long long* dchk = nullptr;
  if (herm_try) {
    use_q = !getenv("SEQ_LH_NOQ") && lhq_plan(N, la, &lq);
    const int nt = (N + LH_T - 1) / LH_T;
    NPh = nt * LH_T;
    std::vector<unsigned> ht;
    size_t packed;      // complex elements of a packed vector
    if (use_q) {
      const int nxt = (lq.L + LQ_T - 1) / LQ_T;
      for (int xi = 0; xi < nxt; xi++)
        for (int xj = xi; xj < nxt; xj++) ht.push_back((unsigned)xi | ((unsigned)xj << 16));
      packed = ht.size() * (size_t)lq.Q * lq.Q * LQ_TE;
      if (lq.Q == 2) { hq_stage = lhq_stage_kernel<2>; hq_first = lhq_first_kernel<2>; hq_pack = lhq_pack_kernel<2>; h_smem = lhq_stage_smem<2>(dfill.nD, dfill.nT); }
      else if (lq.Q == 3) { hq_stage = lhq_stage_kernel<3>; hq_first = lhq_first_kernel<3>; hq_pack = lhq_pack_kernel<3>; h_smem = lhq_stage_smem<3>(dfill.nD, dfill.nT); }
      else { hq_stage = lhq_stage_kernel<4>; hq_first = lhq_first_kernel<4>; hq_pack = lhq_pack_kernel<4>; h_smem = lhq_stage_smem<4>(dfill.nD, dfill.nT); }
      hq_smem_t = (size_t)lq.Q * lq.Q * LQ_T * (LQ_T + 1) * sizeof(cplx);
      if (cudaFuncSetAttribute(hq_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h_smem) != cudaSuccess ||
          cudaFuncSetAttribute(hq_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hq_smem_t) != cudaSuccess ||
          cudaFuncSetAttribute(hq_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hq_smem_t) != cudaSuccess)
        return fail(e, SEQ_ERR_CUDA, "large path: %zu bytes of shared memory per tile refused", h_smem);
    } else {
      for (int ti = 0; ti < nt; ti++)
        for (int tj = ti; tj < nt; tj++) ht.push_back((unsigned)ti | ((unsigned)tj << 16));
      packed = ht.size() * (size_t)LH_TE;
      h_smem = lh_stage_smem(dfill.nD, dfill.nT);
      const char* v = getenv("SEQ_LH_VARIANT");      // experiments: resident CTAs per SM / unrolling of the gather loop
      const int h_variant = v ? atoi(v) : 0;
      h_stage = h_variant == 1 ? lh_stage_kernel<3, 1> : h_variant == 2 ? lh_stage_kernel<2, 2> : lh_stage_kernel<4, 1>;      // (measured at C5: 0.480 / 0.501 / 0.493 ms per 5(4) step for 4 / 3 / 2 x unrolled)
      if (cudaFuncSetAttribute(h_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h_smem) != cudaSuccess)
        return fail(e, SEQ_ERR_CUDA, "large path: %zu bytes of shared memory per tile refused", h_smem);
    }
    h_tiles = (int)ht.size();
    // zero guard rows around the full matrices: gathers are not bounds checked (their coefficients are zero outside)
    int guard = 0;
    for (int t = 0; t < la.nGo; t++) guard = std::max(guard, abs(la.go_off[t]));
    for (int t = 0; t < la.nLg; t++) guard = std::max(guard, std::max(abs(la.lg_oa[t]), abs(la.lg_ob[t])));
    guard += 2 + LQ_T;      // (+ the positions x >= L of the last x tile, which the q-local kernels read but never write)
    const size_t full = ((size_t)NPh + 2 * (size_t)guard) * NPh;      // complex elements
    rc = reserve(e, e->herm, (2 * full + 9 * packed + 8) * sizeof(cplx) + (size_t)h_tiles * sizeof(unsigned) + 256);
    if (rc) return rc;
    cplx* hbase = (cplx*)e->herm.ptr;
    CUDA_TRY(e, cudaMemsetAsync(hbase, 0, 2 * full * sizeof(cplx), st));
    for (int q = 0; q < 2; q++) { hb.Yfull[q] = hbase + (size_t)guard * NPh; hbase += full; }
    for (int q = 0; q < 2; q++) { hb.ybuf[q] = hbase; hbase += packed; }
    for (int q = 0; q < 7; q++) { hb.kbuf[q] = hbase; hbase += packed; }
    dchk = (unsigned long long*)hbase; hbase += 8;
    unsigned* dtiles = (unsigned*)hbase;
    CUDA_TRY(e, cudaMemcpyAsync(dtiles, ht.data(), (size_t)h_tiles * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    hb.tiles = dtiles; hb.NPh = NPh; hb.N = N;
    hb.g0 = dg0; hb.gk = dgk; hb.lam = dlam; hb.gdiag = dgdiag; hb.n_g = n_g; hb.nD = dfill.nD; hb.nT = dfill.nT;
    if (const char* dv = getenv("SEQ_LH_DBG")) {      // timing experiments only (wrong results)
      hb.dbg = atoi(dv);
      if (hb.dbg & 8) { for (int t = 0; t < la.nGo; t++) la.go_off[t] = 0; for (int t = 0; t < la.nLg; t++) la.lg_oa[t] = la.lg_ob[t] = 0; }
      if (hb.dbg & 16) { for (int t = 0; t < la.nGo; t++) la.go_off[t] = (la.go_off[t] > 0) - (la.go_off[t] < 0); for (int t = 0; t < la.nLg; t++) la.lg_oa[t] = la.lg_ob[t] = 1; }
    }
  }
  bool any_full = false;
  for (int b = 0; b < B; b++) {
    bool herm = false;
    if (herm_try) {      // rho0 -> packed upper tiles + full matrix 0, and is it Hermitian?
      CUDA_TRY(e, cudaMemsetAsync(dchk, 0, 2 * sizeof(unsigned long long), st));
      if (use_q) hq_pack<<<h_tiles, 256, hq_smem_t, st>>>(p.state0 + (size_t)b * N * N, hb, lq.L, dchk);
      else lh_pack_kernel<<<h_tiles, 256, 0, st>>>(p.state0 + (size_t)b * N * N, hb, dchk);
      e->launches++;
      double chk[2];
      CUDA_TRY(e, cudaMemcpyAsync(chk, dchk, sizeof(chk), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(e, cudaStreamSynchronize(st));
      herm = chk[1] <= 1e-20 * chk[0];
    }
    if (!herm) {
    // rho0 -> full stage buffer and this rank's rows
    any_full = true;
    large_pack_kernel<<<grid_for(yplane, sms), 256, 0, st>>>(p.state0 + (size_t)b * N * N, Yfull, N, NP, rows_full);
    e->launches++;
    CUDA_TRY(e, cudaMemcpyAsync(ycur, Yfull + row0 * NP, mplane * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(e, cudaMemcpyAsync(ycur + mplane, Yfull + yplane + row0 * NP, mplane * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }

    auto rhs = [&](double t, double* kout) -> int {
      if (n_g > 0) { large_coeff_kernel<<<1, 64, 0, st>>>(p, b, t, cval); e->launches++; }
      if (structured) {
        const int nDN = dfill.nD * N;
        large_diag_gt_kernel<<<(nDN + 255) / 256, 256, 0, st>>>(dg0, dgk, cval, n_g, nDN, dgt);
        LargeDiagArgs a = la;
        a.N = N; a.NP = NP; a.m = m; a.Mrows = Mrows; a.row0 = row0; a.y_plane = yplane; a.k_plane = mplane;
        a.Y = Yfull; a.K = kout; a.gt = dgt; a.lam = dlam; a.gdiag = dgdiag; a.cval = cval;
        large_diag_rhs_kernel<<<grid_for(mplane, sms), 256, 0, st>>>(a);
        e->launches += 2;
        return SEQ_OK;
      }
      large_build_G_kernel<<<grid_for(opsz, sms), 256, 0, st>>>(G0p, Gkp, cval, n_g, opsz, Gt);
      e->launches++;
      CUDA_TRY(e, cudaMemsetAsync(kout, 0, lsz * sizeof(double), st));
      if (Mrows <= 0) return SEQ_OK;
      ZgemmArgs a;
      // K = -i G[R,:] Y
      a.Ar = Gt + row0 * NP; a.Ai = Gt + oplane + row0 * NP; a.lda = NP;
      a.Br = Yfull; a.Bi = Yfull + yplane; a.ldb = NP;
      a.Cr = kout; a.Ci = kout + mplane; a.ldc = NP;
      a.M = Mrows; a.Ncols = NP; a.K = NP; a.alpha_r = 0.0; a.alpha_i = -1.0; a.scale = nullptr;
      a.pref = prefA(0); a.idx = idxA(0);
      int r2 = launch_zgemm<false>(a, sms, st); if (r2) return r2; e->launches++;
      // K += i Y[R,:] G^dag
      a.Ar = Yfull + row0 * NP; a.Ai = Yfull + yplane + row0 * NP;
      a.Br = Gt; a.Bi = Gt + oplane;
      a.alpha_i = 1.0;
      a.pref = prefB(0); a.idx = idxB(0);
      r2 = launch_zgemm<true>(a, sms, st); if (r2) return r2; e->launches++;
      for (int l = 0; l < n_l; l++) {
        const double* Ll = Lp + (size_t)l * opsz;
        CUDA_TRY(e, cudaMemsetAsync(T, 0, lsz * sizeof(double), st));
        a.Ar = Ll + row0 * NP; a.Ai = Ll + oplane + row0 * NP;
        a.Br = Yfull; a.Bi = Yfull + yplane;
        a.Cr = T; a.Ci = T + mplane; a.alpha_r = 1.0; a.alpha_i = 0.0; a.scale = nullptr;
        a.pref = prefA(l + 1); a.idx = idxA(l + 1);
        r2 = launch_zgemm<false>(a, sms, st); if (r2) return r2; e->launches++;
        a.Ar = T; a.Ai = T + mplane;
        a.Br = Ll; a.Bi = Ll + oplane;
        a.Cr = kout; a.Ci = kout + mplane;
        a.scale = (l < p.n_c) ? nullptr : cval + p.n_ch + (l - p.n_c);
        a.pref = prefB(l + 1); a.idx = idxB(l + 1);
        r2 = launch_zgemm<true>(a, sms, st); if (r2) return r2; e->launches++;
      }
      return SEQ_OK;
    };

    StepCtl ctl;
    ctl_init_at(ctl, p, p.t0 + p.dt * (double)out_idx[0]);
    if (devctl) {
      LargeDev h;
      memset(&h, 0, sizeof(h));
      h.ctl = ctl;
      for (int j = 0; j < 7; j++) h.kmap[j] = j;
      CUDA_TRY(e, cudaMemcpyAsync(ddev, &h, sizeof(h), cudaMemcpyHostToDevice, st));
      CUDA_TRY(e, cudaStreamSynchronize(st));      // (h lives on this stack frame)
      LargeDevBufs bf;
      bf.ybuf[0] = ycur; bf.ybuf[1] = ynew;
      for (int j = 0; j < 7; j++) bf.kbuf[j] = kbuf[j];
      bf.Yfull = Yfull; bf.mplane = mplane; bf.yplane = yplane;
      LargeDiagArgs a = la;
      a.N = N; a.NP = NP; a.m = m; a.Mrows = Mrows; a.row0 = row0; a.y_plane = yplane; a.k_plane = mplane;
      a.Y = Yfull; a.K = nullptr; a.gt = dgt; a.lam = dlam; a.gdiag = dgdiag; a.cval = nullptr;
      const int nDN = dfill.nD * N;
      const int g_elem = grid_for(mplane, sms), g_comb = grid_for(lsz, sms);
      auto outputs = [&]() {
        if (n_e > 0) {
          ldev_expect_kernel<<<n_e, 256, 0, st>>>(ddev, e_off, e_ij, e_val, Yfull, yplane, NP, p.expect, (long long)b * n_e * d->n_out, d->n_out,
                                                  (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ? 1 : 0);
          e->launches++;
        }
        if (p.flags & SEQ_FLAG_STORE_STATES) {
          ldev_store_kernel<<<grid_for((long long)N * N, sms), 256, 0, st>>>(ddev, Yfull, p.states + (size_t)b * d->n_out * N * N, N, NP, rows_full);
          e->launches++;
        }
      };
      ldev_ctl_kernel<<<1, 32, 0, st>>>(p, b, ddev, 1);
      e->launches++;
      // (Hermitian tiles: the expectation values of a committed state ride in the next attempt's first kernel)
      LargeHermExpect exa;
      exa.e_off = e_off; exa.e_ij = e_ij; exa.e_val = e_val;
      exa.expect = p.expect; exa.base = (long long)b * n_e * d->n_out; exa.n_out = d->n_out; exa.complex_out = (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ? 1 : 0;
      auto store_h = [&]() {
        if (p.flags & SEQ_FLAG_STORE_STATES) {
          lh_unpack_kernel<<<grid_for((long long)N * N, sms), 256, 0, st>>>(ddev, hb.Yfull[0], p.states + (size_t)b * d->n_out * N * N, (long long)N * N, N, NPh);
          e->launches++;
        }
      };
      if (herm) store_h(); else outputs();
      LargeDev* hdev = (LargeDev*)e->host_err;
      const int chunk = 32;
      long long guard_attempts = 0;
      while (true) {
        for (int q = 0; q < chunk; q++) {
          if (herm) {      // stage input of the first stage, then one fused kernel per stage (RHS + next stage input / error sum)
            if (use_q) launch_pdl(h_pdl, hq_first, h_tiles + n_e, 256, hq_smem_t, st, ddev, hb, lq.L, h_tiles, exa);
            else launch_pdl(h_pdl, lh_first_kernel, h_tiles + n_e, 256, 0, st, ddev, hb, h_tiles, exa);
            e->launches++;
            for (int s_ = 0; s_ < 7; s_++) {
              if (use_q) launch_pdl(h_pdl, hq_stage, h_tiles, 256, h_smem, st, ddev, hb, s_, a, lq, p.atol, p.rtol);
              else launch_pdl(h_pdl, h_stage, h_tiles, 256, h_smem, st, ddev, hb, s_, a, p.atol, p.rtol);
              e->launches++;
            }
            launch_pdl(h_pdl, ldev_ctl_kernel, 1, 32, 0, st, p, b, ddev, 0);
            e->launches++;
            store_h();
            continue;
          }
          for (int s_ = 0; s_ < 7; s_++) {
            if (s_ > 0) { ldev_combine_kernel<<<g_comb, 256, 0, st>>>(ddev, bf, s_); e->launches++; }
            ldev_gt_kernel<<<(nDN + 255) / 256, 256, 0, st>>>(ddev, s_, dg0, dgk, n_g, nDN, dgt);
            e->launches++;
            ldev_rhs_kernel<<<g_elem, 256, 0, st>>>(ddev, bf, s_, a);
            e->launches++;
          }
          ldev_err_kernel<<<g_elem, 256, 0, st>>>(ddev, bf, p.atol, p.rtol, m, NP, N);
          ldev_ctl_kernel<<<1, 32, 0, st>>>(p, b, ddev, 0);
          e->launches += 2;
          outputs();
        }
        CUDA_TRY(e, cudaMemcpyAsync(hdev, ddev, offsetof(LargeDev, cval), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(e, cudaStreamSynchronize(st));
        if (!hdev->active) {
          if (herm && n_e > 0) {      // the last committed state's expectation values, had the chunk ended with its step
            lh_expect_kernel<<<n_e, 256, 0, st>>>(ddev, hb, exa);
            e->launches++;
          }
          break;
        }
        guard_attempts += chunk;
        if (guard_attempts > 100000000LL) return fail(e, SEQ_ERR_CUDA, "large path: the device step controller does not terminate");
      }
      ctl = hdev->ctl;
    } else
    for (int o = 0; o < d->n_out; o++) {
      if (o > 0) {
        const double t_target = p.t0 + p.dt * (double)out_idx[o];
        int nst = 0;
        double htry;
        bool landing, capped;
        while (ctl.status == SEQ_STATUS_OK && ctl_next(ctl, p, t_target, htry, landing, capped)) {
          const int pair = ctl.pair, S = H_RK_STAGES[pair];
          for (int s_ = ctl.have_k1 ? 1 : 0; s_ < S; s_++) {
            if (s_ > 0) {
              LargeCombine c;
              c.y = ycur; c.nk = s_; c.h = htry; c.n_plane = mplane;
              for (int j = 0; j < s_; j++) { c.k[j] = kbuf[j]; c.a[j] = H_RK_A[pair][s_][j]; }
              if (world > 1) { c.dst = Ysend; c.dst_plane = mplane; }
              else { c.dst = Yfull; c.dst_plane = yplane; }
              c.keep = (s_ == S - 1) ? ynew : nullptr;
              large_combine_kernel<<<grid_for(lsz, sms), 256, 0, st>>>(c);
              e->launches++;
              if (world > 1) {
                if (comm->all_gather(comm->ctx, Ysend, Yfull, mplane) || comm->all_gather(comm->ctx, Ysend + mplane, Yfull + yplane, mplane))
                  return fail(e, SEQ_ERR_CUDA, "all_gather hook failed");
              }
            }
            int r2 = rhs(ctl.t + H_RK_C[pair][s_] * htry, kbuf[s_]);
            if (r2) return r2;
            ctl.nrhs++;
          }
          ctl.have_k1 = true;
          LargeErr er;
          er.y = ycur; er.yn = ynew; er.nk = S; er.h = htry; er.atol = p.atol; er.rtol = p.rtol;
          for (int j = 0; j < S; j++) { er.k[j] = kbuf[j]; er.e[j] = H_RK_E[pair][j]; }
          er.m = m; er.NP = NP; er.N = N; er.row0 = row0; er.out = derr;
          CUDA_TRY(e, cudaMemsetAsync(derr, 0, sizeof(double), st));
          large_err_kernel<<<grid_for(mplane, sms), 256, 0, st>>>(er);
          e->launches++;
          if (world > 1 && comm->all_reduce_sum(comm->ctx, derr, 1)) return fail(e, SEQ_ERR_CUDA, "all_reduce_sum hook failed");
          CUDA_TRY(e, cudaMemcpyAsync(e->host_err, derr, sizeof(double), cudaMemcpyDeviceToHost, st));
          CUDA_TRY(e, cudaStreamSynchronize(st));
          const double err = sqrt(e->host_err[0] / ((double)N * N));
          if (ctl_update(ctl, p, err, htry, landing, capped, t_target, nst)) {
            std::swap(ycur, ynew);
            std::swap(kbuf[0], kbuf[S - 1]);
          }
        }
        if (ctl.status != SEQ_STATUS_OK) break;
      }
      // outputs: after an accepted step the full stage buffer holds y (the FSAL stage input)
      if (n_e > 0) {
        CUDA_TRY(e, cudaMemsetAsync(eacc, 0, sizeof(cplx) * n_e, st));
        dim3 ge(grid_for((long long)N * N, sms), n_e);
        large_expect_kernel<<<ge, 256, 0, st>>>(p.E, Yfull, yplane, N, NP, eacc);
        large_store_expect_kernel<<<1, 64, 0, st>>>(eacc, p.expect, (long long)b * n_e * d->n_out + o, d->n_out, n_e,
                                                    (p.flags & SEQ_FLAG_EXPECT_COMPLEX) ? 1 : 0);
        e->launches += 2;
      }
      if (p.flags & SEQ_FLAG_STORE_STATES) {
        large_unpack_kernel<<<grid_for((long long)N * N, sms), 256, 0, st>>>(Yfull, p.states + ((size_t)b * d->n_out + o) * N * N, N, NP, rows_full);
        e->launches++;
      }
    }
    if (ctl.status == SEQ_STATUS_OK) {
      if (herm) lh_unpack_kernel<<<grid_for((long long)N * N, sms), 256, 0, st>>>(nullptr, hb.Yfull[0], p.final_state + (size_t)b * N * N, 0, N, NPh);
      else large_unpack_kernel<<<grid_for((long long)N * N, sms), 256, 0, st>>>(Yfull, p.final_state + (size_t)b * N * N, N, NP, rows_full);
      e->launches++;
    } else {
      CUDA_TRY(e, cudaMemsetAsync(p.final_state + (size_t)b * N * N, 0, sizeof(cplx) * N * N, st));
    }
    int h_stats[4] = {ctl.nacc, ctl.nrej, ctl.nrhs, 0};
    int h_status = ctl.status;
    CUDA_TRY(e, cudaMemcpyAsync(p.status + b, &h_status, sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(e, cudaMemcpyAsync(p.stats + 4 * b, h_stats, sizeof(h_stats), cudaMemcpyHostToDevice, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));   // h_status / h_stats live on this stack frame
  }
  if (herm_try && !any_full) {
    char vb[256];
    if (use_q) snprintf(vb, sizeof(vb), "large/diagonals, Hermitian tiles, outer factor of dimension %d inside a thread (device-resident step control, stage combination and error estimate fused into the RHS kernel)", lq.Q);
    else snprintf(vb, sizeof(vb), "large/diagonals, Hermitian tiles (device-resident step control, stage combination and error estimate fused into the RHS kernel)");
    e->variant = vb;
  }
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

static int launch_sparse(seq_engine* e, const seq_solve_desc* d, SolveParams& p, const cplx* G0, const cplx* Gk, const SparseInfo& info) {
  const int N = d->N, n_g = p.n_g, n_l = p.n_c + p.n_ctd, n_e = p.n_e;
  const SparsePlan& pl = info.plan;
  SparsePack pk = sparse_pack_layout(N, n_g, n_l, n_e, pl);
  int rc = reserve(e, e->pack, pk.bytes);
  if (rc) return rc;
  rc = reserve(e, e->ws, sparse_ws_elems(pl) * (size_t)d->B * sizeof(cplx) + 256);
  if (rc) return rc;
  char* base = (char*)e->pack.ptr;
  SparseOps sp;
  sp.wG = pl.wG; sp.wL = pl.wL; sp.n_el = N * (N + 1) / 2; sp.ld = pl.ld; sp.ybufs = pl.ybufs;
  sp.gcol = (int*)(base + pk.o_gcol); sp.g0 = (cplx*)(base + pk.o_g0); sp.gk = (cplx*)(base + pk.o_gk);
  sp.lcol = (int*)(base + pk.o_lcol); sp.lval = (cplx*)(base + pk.o_lval); sp.ij = (unsigned*)(base + pk.o_ij);
  sp.e_off = (int*)(base + pk.o_eoff); sp.e_ij = (unsigned*)(base + pk.o_eij); sp.e_val = (cplx*)(base + pk.o_eval);
  const int n_ops = 1 + n_l + n_e;
  int* rowoff = info.rowcnt + (size_t)n_ops * N + 2 * (size_t)n_ops;
  sparse_escan_kernel<<<1, 32, 0, e->stream>>>(info.rowcnt, n_l, n_e, N, rowoff, (int*)sp.e_off);
  dim3 grid(N, n_ops);
  sparse_fill_kernel<<<grid, 32, 0, e->stream>>>(G0, Gk, n_g, p.L, n_l, p.E, N, pl.wG, pl.wL, (int*)sp.gcol, (cplx*)sp.g0, (cplx*)sp.gk,
                                                   (int*)sp.lcol, (cplx*)sp.lval, rowoff, (unsigned*)sp.e_ij, (cplx*)sp.e_val, 0);
  sparse_ij_kernel<<<(N * N + 255) / 256, 256, 0, e->stream>>>((unsigned*)sp.ij, N);
  e->launches += 3;
  p.ws = (cplx*)e->ws.ptr;
  p.ws_stride = (long long)sparse_ws_elems(pl);
  CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  rc = launch_sparse_kernel(p, sp, pl, e->stream);
  if (rc) return fail(e, rc, "sparse kernel launch configuration failed for N=%d", N);
  e->launches++;
  return SEQ_OK;
}

// Diagonal structure of the operators (device flags -> host plan).  One D2H copy + synchronisation.
// `tail`: a second plan (2-CTA clusters per trajectory) for the last partial wave of a launch, see launch_diag
static int diag_analyze(seq_engine* e, const seq_solve_desc* d, const cplx* G0, const cplx* Gk, const cplx* Lall, int nnzE, DiagPlan* pl, bool* usable,
                        DiagPlan* tail, bool* tail_ok, BlkPlan* blk, bool* blk_ok) {
  const int N = d->N, n_g = d->n_ch + d->n_ctd, n_l = d->n_c + d->n_ctd;
  const size_t W = 2 * (size_t)N - 1, nflag = (size_t)(2 + n_l) * W;
  *usable = false;
  int rc = reserve(e, e->dflags, nflag * sizeof(int) + 64);
  if (rc) return rc;
  int* flags = (int*)e->dflags.ptr;
  CUDA_TRY(e, cudaMemsetAsync(flags, 0, nflag * sizeof(int), e->stream));
  dim3 grid((unsigned)grid_for((long long)N * N, e->sm_count), 2 + n_l);
  diag_flags_kernel<<<grid, 256, 0, e->stream>>>(G0, Gk, n_g, Lall, N, flags);
  e->launches++;
  std::vector<int> h(nflag);
  CUDA_TRY(e, cudaMemcpyAsync(h.data(), flags, nflag * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  const char* env = getenv("SEQ_DIAG_CL");
  const char* cap = getenv("SEQ_DIAG_SMEM_CAP");
  // SEQ_DIAG_CW=0: no controller warp (N <= 48 runs with one by default: 5 % faster on trajectories that stay on the
  // capped 4(3) pair, 4 % slower on those that keep returning to the 5(4) pair - the 13th warp costs the compute
  // warps 40 registers; 3 % faster on the C2 sweep), see DESIGN.md 4.1b
  const char* cw = getenv("SEQ_DIAG_CW");
  int max_smem = e->max_smem_optin;
  if (cap && atoi(cap) > 0 && atoi(cap) < max_smem) max_smem = atoi(cap);
  *usable = diag_plan(N, n_g, d->n_c, d->n_ctd, d->n_ch, d->n_t, nnzE, h.data(), max_smem, env ? atoi(env) : 0, !(cw && atoi(cw) == 0), pl);
  *tail_ok = false;
  // block-local variant (kernel_diagblk.cuh): is there a small tensor factor whose operators can stay inside a thread?
  // Candidates come from the diagonal offsets; which diagonals really are local to a candidate is checked on the data.
  *blk_ok = false;
  const char* noblk = getenv("SEQ_DIAG_BLK");
  if (*usable && !(noblk && atoi(noblk) == 0) && !(env && atoi(env) == 2)) {
    BlkCands cands;
    diagblk_candidates(N, n_l, h.data(), &cands);
    if (cands.n > 0) {
      const size_t ncross = 2 * (size_t)cands.n * (1 + n_l) * W;      // crossing flags, then local-index-change flags
      rc = reserve(e, e->dflags, (nflag + ncross) * sizeof(int) + 64);
      if (rc) return rc;
      int* cross = (int*)e->dflags.ptr + nflag;
      CUDA_TRY(e, cudaMemsetAsync(cross, 0, ncross * sizeof(int), e->stream));
      dim3 gc((unsigned)grid_for((long long)N * N, e->sm_count), 1 + n_l, cands.n);
      diagblk_cross_kernel<<<gc, 256, 0, e->stream>>>(cands, G0, Gk, n_g, Lall, N, cross);
      e->launches++;
      std::vector<int> hc(ncross);
      CUDA_TRY(e, cudaMemcpyAsync(hc.data(), cross, ncross * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      BlkPlan cand;
      for (int c = 0; c < cands.n; c++) {
        if (!diagblk_plan(N, n_g, d->n_c, d->n_ctd, d->n_ch, d->n_t, nnzE, h.data(), hc.data() + (size_t)c * (1 + n_l) * W,
                          hc.data() + (size_t)(cands.n + c) * (1 + n_l) * W, cands.s[c], cands.D[c], max_smem,
                          e->sm_count, &cand))
          continue;
        if (cand.gathers >= cand.gathers_plain) continue;      // nothing became local
        if (!*blk_ok || cand.gathers < blk->gathers || (cand.gathers == blk->gathers && cand.bl.D > blk->bl.D)) { *blk = cand; *blk_ok = true; }
      }
    }
  }
  // A batch that does not even fill one CTA per SM is latency-bound per trajectory, and there the plain kernel's twelve warps
  // on one trajectory finish it in 0.72 of the time the block-local kernel's four do (which wins by running two trajectories
  // per SM).  Under AUTO such batches keep the plain kernel - e.g. a 1024-point sweep strong-scaled over 8 GPUs.
  if (*blk_ok && blk->minb > 1 && d->method == SEQ_METHOD_AUTO && d->B <= e->sm_count) *blk_ok = false;
  const int rest = d->B % e->sm_count;
  if (*usable && pl->cl == 1 && !e->is_sub && !getenv("SEQ_DIAG_NO_TAIL") && d->B > e->sm_count && rest > 0 && 2 * rest <= e->sm_count && N >= 16)
    *tail_ok = diag_plan(N, n_g, d->n_c, d->n_ctd, d->n_ch, d->n_t, nnzE, h.data(), max_smem, 2, false, tail) && tail->cl == 2 && tail->kreg;
  return SEQ_OK;
}

// One CTA per trajectory and one CTA per SM: a batch that is not a multiple of the SM count ends in a partial wave
// during which most SMs idle (C4: 512 trajectories = 3.46 waves on 148 SMs).  When that rest is at most half a wave
// its trajectories run on 2-CTA clusters instead (`tail`): both halves of the GPU work on them and the wave takes
// ~0.63 of the time.  The cluster variant sums the error norm in a different order, so a tail trajectory agrees with
// the others to rounding (1e-13), not bit for bit; SEQ_DIAG_NO_TAIL=1 switches the split off.
static int launch_diag(seq_engine* e, const seq_solve_desc* d, SolveParams& p, const cplx* G0, const cplx* Gk, const SparseInfo& info, DiagPlan& pl_plain,
                       DiagPlan* tail, BlkPlan* blk, bool plain_ok = true) {
  DiagPlan& pl = blk ? blk->dp : pl_plain;
  if (blk && blk->minb > 1) tail = nullptr;      // (two CTAs per SM: the partial last wave goes to the plain kernel instead, below)
  const int N = d->N, n_g = p.n_g, n_l = p.n_c + p.n_ctd, n_e = p.n_e;
  // expectation operators as COO lists (layout shared with the ELL kernel), operator diagonals behind them
  SparsePlan epl;
  memset(&epl, 0, sizeof(epl));
  epl.wG = 1; epl.wL = 1; epl.nnzE = info.plan.nnzE;
  SparsePack pk = sparse_pack_layout(N, 0, 0, n_e, epl);
  DiagPack dk = diag_pack_layout(N, n_g, pl);
  int rc = reserve(e, e->pack, pk.bytes + dk.bytes);
  if (rc) return rc;
  const size_t ws_per = blk ? diagblk_ws_elems(*blk) : diag_ws_elems(pl), ws_shared = blk ? diagblk_ws_shared_elems(*blk) : diag_ws_shared_elems(pl);
  rc = reserve(e, e->ws, (ws_per * (size_t)d->B + ws_shared) * sizeof(cplx) + 256);
  if (rc) return rc;
  char* base = (char*)e->pack.ptr;
  char* dbase = base + pk.bytes / 16 * 16;
  DiagOps& sp = pl.ops;
  sp.e_off = (int*)(base + pk.o_eoff); sp.e_ij = (unsigned*)(base + pk.o_eij); sp.e_val = (cplx*)(base + pk.o_eval);
  sp.g0 = (cplx*)(dbase + dk.o_g0); sp.gk = (cplx*)(dbase + dk.o_gk); sp.lam = (cplx*)(dbase + dk.o_lam);
  sp.gdiag = (cplx*)(dbase + dk.o_gdiag);
  int* any_imag = (int*)(dbase + dk.o_flag);
  const int n_ops = 1 + n_l + n_e;
  int* rowoff = info.rowcnt + (size_t)n_ops * N + 2 * (size_t)n_ops;
  sparse_escan_kernel<<<1, 32, 0, e->stream>>>(info.rowcnt, n_l, n_e, N, rowoff, (int*)sp.e_off);
  e->launches++;
  if (n_e > 0) {
    dim3 ge(N, n_e);
    sparse_fill_kernel<<<ge, 32, 0, e->stream>>>(G0, Gk, n_g, p.L, n_l, p.E, N, 1, 1, nullptr, nullptr, nullptr, nullptr, nullptr, rowoff,
                                               (unsigned*)sp.e_ij, (cplx*)sp.e_val, 1 + n_l);
    e->launches++;
  }
  CUDA_TRY(e, cudaMemsetAsync(any_imag, 0, sizeof(int), e->stream));
  dim3 gf((N + 127) / 128, pl.fill.nD + pl.fill.nT);
  diag_fill_kernel<<<gf, 128, 0, e->stream>>>(pl.fill, G0, Gk, n_g, p.L, N, (cplx*)sp.g0, (cplx*)sp.gk, (cplx*)sp.lam, (cplx*)sp.gdiag, any_imag);
  e->launches++;
  int h_imag = 0;
  CUDA_TRY(e, cudaMemcpyAsync(&h_imag, any_imag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  p.ws = (cplx*)e->ws.ptr;
  p.ws_stride = (long long)ws_per;
  CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  sp.lreal = h_imag == 0 ? 1 : 0;
  if (blk) {
    const char* why_mirror = blk->bl.mirror ? ", mirrored blocks written (a gather leaves the owned band)" : "";
    if (p.flags & SEQ_FLAG_STORE_STATES) { blk->bl.mirror = 1; why_mirror = ", mirrored blocks written (stored states read the full matrix)"; }
    else if (!sp.stage_e && sp.nnzE > 0) { blk->bl.mirror = 1; why_mirror = ", mirrored blocks written (expectation lists not staged)"; }
    char buf[448];
    snprintf(buf, sizeof(buf), "diagblk: local mode dim %d stride %d kept in registers, 1 CTA x %d threads per trajectory (%d CTA%s/SM), %d elements/thread, "
             "rolled stage loop, %s%d stage-derivative slots in shared memory + %d in L2, %d Y buffer%s%s%s, %d gathers/element (plain diag: %d), smem %zu B",
             blk->bl.D, blk->bl.s, blk->threads, blk->minb, blk->minb > 1 ? "s" : "", blk->bl.D * blk->bl.D, blk->k1reg ? "k1 in registers, " : "", blk->bl.kslots, blk->gslots,
             sp.ybufs, sp.ybufs > 1 ? "s" : "", pl.guard ? " + guard bands" : "", blk->bl.mirror ? why_mirror : ", owned half only (no mirror writes)", blk->gathers,
             blk->gathers_plain, pl.smem);
    e->variant = buf;
    e->last_fma_per_rhs = blk->fma_per_rhs;
    e->last_fp64_step_extra = blk->fp64_step_extra;
    // Partial last wave.  Two CTAs per SM integrate two trajectories in the time one CTA alone needs for one (the kernel is
    // latency-bound per CTA), so a batch that ends in at most half a wave of CTA slots leaves every SM half empty for a whole
    // trajectory time (C2: 1024 = 3 x 296 + 136).  The plain diagonal kernel - 12 warps on ONE trajectory, one CTA per SM -
    // finishes a trajectory in 0.72 of that time when it has the SM to itself: the rest runs on it, on a second stream, next
    // to the main launch.  Same equations and step decisions; the two kernels order their sums differently, so a tail
    // trajectory agrees with the block-local kernel to rounding (1e-15), not bit for bit.  SEQ_DIAG_NO_TAIL=1 switches it off.
    const long long slots = (long long)blk->minb * e->sm_count, rest = d->B % slots;
    const bool plain_tail = blk->minb > 1 && !getenv("SEQ_DIAG_NO_TAIL") && d->B > slots && rest > 0 && rest <= e->sm_count &&
                            plain_ok && pl_plain.cl == 1 && pl_plain.kreg;
    if (plain_tail) {
      const long long b1 = d->B - rest;
      SolveParams p1 = p, p2 = p;
      p1.B = (int)b1;
      p2.B = (int)rest;
      p2.tab = p.tab + b1 * p.tab_bstride_c;
      p2.tab_rate = p.tab_rate + b1 * p.tab_bstride_r;
      if (p.coeff_scale) p2.coeff_scale = p.coeff_scale + b1 * p.n_ch;
      p2.state0 = p.state0 + b1 * N * N;
      if (p.expect) p2.expect = (char*)p.expect + (size_t)b1 * p.n_e * p.n_out * ((p.flags & SEQ_FLAG_EXPECT_COMPLEX) ? sizeof(cplx) : sizeof(double));
      p2.final_state = p.final_state + b1 * N * N;
      if (p.states) p2.states = p.states + b1 * (long long)p.n_out * N * N;
      p2.status = p.status + b1;
      p2.stats = p.stats + 4 * b1;
      p2.ws_stride = 0;
      DiagOps& sp2 = pl_plain.ops;      // same tables: both plans walk the diagonals in the same order
      sp2.e_off = sp.e_off; sp2.e_ij = sp.e_ij; sp2.e_val = sp.e_val;
      sp2.g0 = sp.g0; sp2.gk = sp.gk; sp2.lam = sp.lam; sp2.gdiag = sp.gdiag; sp2.lreal = sp.lreal;
      if (!e->aux) {
        CUDA_TRY(e, cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
        CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_aux0, cudaEventDisableTiming));
        CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_aux1, cudaEventDisableTiming));
      }
      CUDA_TRY(e, cudaEventRecord(e->ev_aux0, e->stream));
      CUDA_TRY(e, cudaStreamWaitEvent(e->aux, e->ev_aux0, 0));
      // (the main launch first: its CTAs fill whole waves of CTA slots; the tail's CTAs - one per SM, they need the whole
      // shared memory - start as SMs drain at the end of the last full wave)
      rc = launch_diagblk_kernel(p1, *blk, e->stream);
      if (rc == SEQ_OK) rc = launch_diag_kernel(p2, sp2, pl_plain, e->aux);
      if (rc) return fail(e, rc, "block-local diagonal kernel launch configuration failed for N=%d (tail split)", N);
      CUDA_TRY(e, cudaEventRecord(e->ev_aux1, e->aux));
      CUDA_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_aux1, 0));
      e->launches += 2;
      char b2[128];
      snprintf(b2, sizeof(b2), "; last %lld trajectories on the plain diagonal kernel (%d threads%s, 1 CTA/SM)", rest, pl_plain.threads, pl_plain.cw ? " + controller warp" : "");
      e->variant += b2;
      return SEQ_OK;
    }
    if (!tail) {
      rc = launch_diagblk_kernel(p, *blk, e->stream);
      if (rc) return fail(e, rc, "block-local diagonal kernel launch configuration failed for N=%d", N);
      e->launches++;
      return SEQ_OK;
    }
    // (one CTA per SM: a rest of at most half a wave runs on the plain kernel's 2-CTA clusters, as it does for the plain kernel)
  } else {
    char buf[224];
    snprintf(buf, sizeof(buf), "diag: %d CTA%s x %d threads%s per trajectory, %d elements/thread, stage vectors in %s, %d Y buffer%s%s, smem %zu B", pl.cl, pl.cl > 1 ? "s (cluster)" : "",
             pl.threads, pl.cw ? (sp.ybufs > 1 ? " + controller warp (speculative next step)" : " + controller warp") : "", pl.ept, pl.kreg ? "registers" : "L2 streams", sp.ybufs,
             sp.ybufs > 1 ? "s" : "", pl.guard ? " + guard bands" : "", pl.smem);
    e->variant = buf;
  }
  if (tail) {
    const long long nt = d->B % e->sm_count, b1 = d->B - nt;
    SolveParams p1 = p, p2 = p;
    p1.B = (int)b1;
    p2.B = (int)nt;
    p2.tab = p.tab + b1 * p.tab_bstride_c;
    p2.tab_rate = p.tab_rate + b1 * p.tab_bstride_r;
    if (p.coeff_scale) p2.coeff_scale = p.coeff_scale + b1 * p.n_ch;
    p2.state0 = p.state0 + b1 * N * N;
    if (p.expect) p2.expect = (char*)p.expect + (size_t)b1 * p.n_e * p.n_out * ((p.flags & SEQ_FLAG_EXPECT_COMPLEX) ? sizeof(cplx) : sizeof(double));
    p2.final_state = p.final_state + b1 * N * N;
    if (p.states) p2.states = p.states + b1 * (long long)p.n_out * N * N;
    p2.status = p.status + b1;
    p2.stats = p.stats + 4 * b1;
    p2.ws_stride = 0;
    DiagOps& sp2 = tail->ops;
    sp2.e_off = sp.e_off; sp2.e_ij = sp.e_ij; sp2.e_val = sp.e_val;
    sp2.g0 = sp.g0; sp2.gk = sp.gk; sp2.lam = sp.lam; sp2.gdiag = sp.gdiag; sp2.lreal = sp.lreal;
    // the clusters go first, on a second stream, so that they find pairs of free SMs; the main launch fills the rest of
    // the GPU and both drain together (one stream would put a grid-wide barrier between the main launch and the tail)
    if (!e->aux) {
      CUDA_TRY(e, cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
      CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_aux0, cudaEventDisableTiming));
      CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_aux1, cudaEventDisableTiming));
    }
    CUDA_TRY(e, cudaEventRecord(e->ev_aux0, e->stream));
    CUDA_TRY(e, cudaStreamWaitEvent(e->aux, e->ev_aux0, 0));
    rc = launch_diag_kernel(p2, sp2, *tail, e->aux);
    if (rc == SEQ_OK) rc = blk ? launch_diagblk_kernel(p1, *blk, e->stream) : launch_diag_kernel(p1, sp, pl, e->stream);
    if (rc) return fail(e, rc, "diagonal kernel launch configuration failed for N=%d (tail split)", N);
    CUDA_TRY(e, cudaEventRecord(e->ev_aux1, e->aux));
    CUDA_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_aux1, 0));
    e->launches += 2;
    char buf[96];
    snprintf(buf, sizeof(buf), "; last %lld trajectories on 2-CTA clusters x %d threads", nt, tail->threads);
    e->variant += buf;
    return SEQ_OK;
  }
  rc = launch_diag_kernel(p, sp, pl, e->stream);
  if (rc) return fail(e, rc, "diagonal kernel launch configuration failed for N=%d", N);
  e->launches++;
  return SEQ_OK;
}

// Kets (sesolve path): operators by diagonals when they have that structure (kernel_ket.cuh).  `forced`: SEQ_METHOD_DIAG was
// asked for.  *ran = false: the operators are not diagonal-structured enough - the caller falls back to the generic kernel.
static int launch_ket(seq_engine* e, const seq_solve_desc* d, SolveParams& p, const cplx* G0, const cplx* Gk, bool forced, bool* ran) {
  const int N = d->N, n_ch = d->n_ch, n_e = d->n_e;
  *ran = false;
  if (d->n_c + d->n_ctd > 0 || N > 0x7fff) return SEQ_OK;
  const size_t W = 2 * (size_t)N - 1, nflag = (size_t)(1 + n_ch) * W;
  int rc = reserve(e, e->dflags, nflag * sizeof(int) + 64);
  if (rc) return rc;
  int* flags = (int*)e->dflags.ptr;
  CUDA_TRY(e, cudaMemsetAsync(flags, 0, nflag * sizeof(int), e->stream));
  dim3 grid((unsigned)grid_for((long long)N * N, e->sm_count), 1 + n_ch);
  ket_flags_kernel<<<grid, 256, 0, e->stream>>>(G0, Gk, N, flags);
  e->launches++;
  std::vector<int> h(nflag);
  CUDA_TRY(e, cudaMemcpyAsync(h.data(), flags, nflag * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SparseInfo spi;
  rc = sparse_analyze(e, d, G0, Gk, nullptr, &spi);      // (synchronises; only the expectation operators' row counts are used)
  if (rc) return rc;
  KetPlan pl;
  if (!ket_plan(N, n_ch, p.n_g, spi.plan.nnzE, h.data(), e->max_smem_optin, &pl)) return SEQ_OK;
  if (!forced && pl.macs * 4 > (long long)N) return SEQ_OK;      // not sparse enough to beat the dense product
  SparsePlan epl;
  memset(&epl, 0, sizeof(epl));
  epl.wG = 1; epl.wL = 1; epl.nnzE = spi.plan.nnzE;
  SparsePack pk = sparse_pack_layout(N, 0, 0, n_e, epl);
  const size_t tab_bytes = (size_t)pl.ops.nTab * N * sizeof(cplx);
  rc = reserve(e, e->pack, pk.bytes + tab_bytes + 256);
  if (rc) return rc;
  char* base = (char*)e->pack.ptr;
  KetOps& ko = pl.ops;
  ko.e_off = (int*)(base + pk.o_eoff); ko.e_ij = (unsigned*)(base + pk.o_eij); ko.e_val = (cplx*)(base + pk.o_eval);
  ko.tabs = (cplx*)(base + (pk.bytes + 15) / 16 * 16);
  const int n_ops = 1 + n_e;
  int* rowoff = spi.rowcnt + (size_t)n_ops * N + 2 * (size_t)n_ops;
  sparse_escan_kernel<<<1, 32, 0, e->stream>>>(spi.rowcnt, 0, n_e, N, rowoff, (int*)ko.e_off);
  e->launches++;
  if (n_e > 0) {
    dim3 ge(N, n_e);
    sparse_fill_kernel<<<ge, 32, 0, e->stream>>>(G0, Gk, p.n_g, p.L, 0, p.E, N, 1, 1, nullptr, nullptr, nullptr, nullptr, nullptr, rowoff,
                                               (unsigned*)ko.e_ij, (cplx*)ko.e_val, 1);
    e->launches++;
  }
  dim3 gf((N + 127) / 128, pl.fill.nTab);
  ket_fill_kernel<<<gf, 128, 0, e->stream>>>(pl.fill, G0, Gk, N, (cplx*)ko.tabs);
  e->launches++;
  CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  char buf[256];
  snprintf(buf, sizeof(buf), "ketdiag: operators by diagonals, %d terms per element (dense product: %d), 1 CTA x %d threads per trajectory, %d element%s/thread, "
           "all vectors in shared memory, smem %zu B", pl.ops.nTerm + pl.ops.has_w, N, pl.threads, pl.ept, pl.ept > 1 ? "s" : "", pl.smem);
  e->variant = buf;
  e->last_fma_per_rhs = 4LL * (pl.ops.nTerm + pl.ops.has_w) * N;
  rc = launch_ket_kernel(p, pl, e->stream);
  if (rc) return fail(e, rc, "ket kernel launch configuration failed for N=%d", N);
  e->launches++;
  *ran = true;
  return SEQ_OK;
}

static int solve_device(seq_engine* e, const seq_solve_desc* d) {
  const int N = d->N, B = d->B;
  const size_t N2 = (size_t)N * N;
  const int n_g = d->n_ch + d->n_ctd;
  const bool ket = d->flags & SEQ_FLAG_KET;
  int method = pick_method(e, d);
  const bool sharded = e->comm && e->comm->world > 1;
  // kets: the operators-by-diagonals kernel is tried whenever the choice is the engine's (or SEQ_METHOD_DIAG is asked for);
  // SEQ_KET_DIAG=0 keeps the dense generic kernel
  const char* kd = getenv("SEQ_KET_DIAG");
  const bool ket_diag_wanted = ket && !d->H0_batch_stride && !sharded && (d->method == SEQ_METHOD_AUTO || d->method == SEQ_METHOD_DIAG) &&
                               !(kd && atoi(kd) == 0 && d->method != SEQ_METHOD_DIAG);
  if (ket && d->method == SEQ_METHOD_DIAG) {
    if (!ket_diag_wanted) return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_DIAG on kets needs a shared H0 on one GPU");
    method = SEQ_METHOD_GENERIC;      // (scratch sizing; launch_ket runs first and fails loudly if it cannot)
  }
  // the sparse kernel is tried for density matrices with a shared H0 whenever the choice is the engine's
  const bool try_sparse = !ket && !d->H0_batch_stride && !sharded &&
                          (method == SEQ_METHOD_SPARSE || method == SEQ_METHOD_DIAG ||
                           (d->method == SEQ_METHOD_AUTO && method != SEQ_METHOD_SMALL));
  if ((method == SEQ_METHOD_SPARSE || method == SEQ_METHOD_DIAG) && !try_sparse)
    return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_SPARSE / SEQ_METHOD_DIAG cover density matrices with a shared H0 on one GPU (N=%d ket=%d)", N, (int)ket);
  if (method == SEQ_METHOD_DMMA && (ket || !dmma_supported(N, e->max_smem_optin) || d->H0_batch_stride))
    return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_DMMA does not cover N=%d ket=%d h0_stride=%lld", N, (int)ket, (long long)d->H0_batch_stride);
  if (method == SEQ_METHOD_SMALL && (ket || d->H0_batch_stride || !small_supported(N, n_g, d->n_e)))
    return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_SMALL does not cover N=%d ket=%d h0_stride=%lld (density matrices, N <= 6, shared H0)", N, (int)ket, (long long)d->H0_batch_stride);
  if (method == SEQ_METHOD_DMMA_STREAM && (ket || !stream_supported(N, e->max_smem_optin) || d->H0_batch_stride))
    return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_DMMA_STREAM does not cover N=%d ket=%d h0_stride=%lld", N, (int)ket, (long long)d->H0_batch_stride);
  if (method == SEQ_METHOD_LARGE && (ket || d->H0_batch_stride))
    return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_LARGE covers density matrices with a shared H0 only (N=%d ket=%d)", N, (int)ket);
  if (method < SEQ_METHOD_GENERIC || method > SEQ_METHOD_DIAG)
    return fail(e, SEQ_ERR_UNSUPPORTED, "method %d is not implemented", method);
  if (sharded && method != SEQ_METHOD_LARGE)
    return fail(e, SEQ_ERR_UNSUPPORTED, "row-sharded solves run SEQ_METHOD_LARGE only (got method %d for N=%d)", method, N);

  // ---- scratch
  const size_t nG0 = d->H0_batch_stride ? (size_t)B : 1;
  int rc = reserve(e, e->gbuf, (nG0 + n_g) * N2 * sizeof(cplx));
  if (rc) return rc;
  const long long series_c = (long long)(d->coeff_batch_stride ? B : 1) * d->n_ch;
  const long long series_r = (long long)(d->rate_batch_stride ? B : 1) * d->n_ctd;
  rc = reserve(e, e->tab, (size_t)(series_c + series_r) * d->n_t * sizeof(double2) + 256);
  if (rc) return rc;
  if (d->coeff_batch_stride && d->coeff_batch_stride != (int64_t)d->n_ch * d->n_t)
    return fail(e, SEQ_ERR_UNSUPPORTED, "coeff_batch_stride must be 0 or n_ch*n_t");
  if (d->rate_batch_stride && d->rate_batch_stride != (int64_t)d->n_ctd * d->n_t)
    return fail(e, SEQ_ERR_UNSUPPORTED, "rate_batch_stride must be 0 or n_ctd*n_t");

  cplx* G0 = (cplx*)e->gbuf.ptr;
  cplx* Gk = G0 + nG0 * N2;
  double2* tabc = (double2*)e->tab.ptr;
  double2* tabr = tabc + (size_t)series_c * d->n_t;

  // ---- preparation
  {
    const bool rows_prep = N > 128 && !ket && nG0 == 1;      // large N: column-driven L^dag L (see prep_ldagl_rows_kernel)
    if (rows_prep) {
      prep_ldagl_rows_kernel<<<dim3(N, 1), 128, 2 * N * sizeof(int), e->stream>>>((const cplx*)d->H0, (const cplx*)d->Lc, d->n_c, 0, N, G0);
      e->launches++;
      if (d->n_ch > 0) CUDA_TRY(e, cudaMemcpyAsync(Gk, d->Hk, (size_t)d->n_ch * N2 * sizeof(cplx), cudaMemcpyDeviceToDevice, e->stream));
      if (d->n_ctd > 0) {
        prep_ldagl_rows_kernel<<<dim3(N, d->n_ctd), 128, 2 * N * sizeof(int), e->stream>>>(nullptr, (const cplx*)d->Ltd, d->n_ctd, 1, N, Gk + (size_t)d->n_ch * N2);
        e->launches++;
      }
    } else {
    dim3 grid((unsigned)((N2 + 127) / 128), (unsigned)nG0);
    prep_G0_kernel<<<grid, 128, 0, e->stream>>>((const cplx*)d->H0, d->H0_batch_stride, (const cplx*)d->Lc, d->n_c, N, G0, ket ? 1 : 0);
    e->launches++;
    if (n_g > 0) {
      dim3 gk((unsigned)((N2 + 127) / 128), (unsigned)n_g);
      prep_Gk_kernel<<<gk, 128, 0, e->stream>>>((const cplx*)d->Hk, d->n_ch, (const cplx*)d->Ltd, d->n_ctd, N, Gk);
      e->launches++;
    }
    }
    rc = spline_prep_packed(e, d->coeff, tabc, series_c, d->n_t);
    if (rc) return rc;
    rc = spline_prep_packed(e, d->rate, tabr, series_r, d->n_t);
    if (rc) return rc;
  }

  // collapse operators must be contiguous [n_c + n_ctd, N, N] for the kernels: if both kinds
  // are present, pack them behind the G matrices
  const cplx* Lall = (const cplx*)d->Lc;
  if (d->n_ctd > 0) {
    if (d->n_c == 0) {
      Lall = (const cplx*)d->Ltd;
    } else {
      size_t need = (nG0 + n_g + d->n_c + d->n_ctd) * N2 * sizeof(cplx);
      // gbuf may move: re-reserve before the prep launches would be cleaner, but reserve() syncs the stream first
      if (need > e->gbuf.cap) return fail(e, SEQ_ERR_UNSUPPORTED, "internal: gbuf too small for packed collapse operators");
      cplx* dst = Gk + (size_t)n_g * N2;
      CUDA_TRY(e, cudaMemcpyAsync(dst, d->Lc, (size_t)d->n_c * N2 * sizeof(cplx), cudaMemcpyDeviceToDevice, e->stream));
      CUDA_TRY(e, cudaMemcpyAsync(dst + (size_t)d->n_c * N2, d->Ltd, (size_t)d->n_ctd * N2 * sizeof(cplx), cudaMemcpyDeviceToDevice, e->stream));
      Lall = dst;
    }
  }

  // ---- operator sparsity decides between the sparse (FMA) and the dense (tensor-core) formulation
  SparseInfo spi;
  DiagPlan dpl, dpl_tail;
  BlkPlan bpl;
  bool tail_ok = false, blk_ok = false;
  if (try_sparse) {
    rc = sparse_analyze(e, d, G0, Gk, Lall, &spi);
    if (rc) return rc;
    bool diag_ok = false;
    if (method != SEQ_METHOD_SPARSE) {
      rc = diag_analyze(e, d, G0, Gk, Lall, spi.plan.nnzE, &dpl, &diag_ok, &dpl_tail, &tail_ok, &bpl, &blk_ok);
      if (rc) return rc;
    }
    if (method == SEQ_METHOD_SPARSE) {
      if (!spi.usable) return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_SPARSE: operator rows too dense or N=%d too large for one CTA", N);
    } else if (method == SEQ_METHOD_DIAG) {
      if (!diag_ok) return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_DIAG: operators have too many diagonals or N=%d too large for one CTA", N);
    } else {
      const int n_l = d->n_c + d->n_ctd;
      const bool diag_worth = diag_ok && dpl.macs * 6 <= (long long)N * (1 + 2 * n_l);
      if (diag_worth && (!(spi.usable && spi.worthwhile) || dpl.macs <= 2 * spi.plan.macs)) method = SEQ_METHOD_DIAG;
      else if (spi.usable && spi.worthwhile) method = SEQ_METHOD_SPARSE;
    }
  }
  e->last_method = method;
  e->last_fma_per_rhs = 0;
  e->last_fp64_step_extra = 0;
  e->variant.clear();

  const size_t ws_per = ws_elems_per_traj(d, method);
  rc = reserve(e, e->ws, ws_per * (size_t)B * sizeof(cplx) + 256);
  if (rc) return rc;
  if (method == SEQ_METHOD_DMMA) {
    rc = reserve(e, e->pack, dmma_pack_doubles(N, n_g, d->n_c + d->n_ctd, d->n_e) * sizeof(double));
    if (rc) return rc;
  }
  if (method == SEQ_METHOD_DMMA_STREAM) {
    rc = reserve(e, e->pack, stream_pack_doubles(N, n_g, d->n_c + d->n_ctd, d->n_e) * sizeof(double));
    if (rc) return rc;
  }
  if (method == SEQ_METHOD_SMALL) {
    rc = reserve(e, e->pack, small_pack_doubles(N, n_g, d->n_e) * sizeof(double));
    if (rc) return rc;
  }
  SolveParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.B = B; p.n_t = d->n_t; p.n_g = n_g; p.n_ch = d->n_ch; p.n_c = d->n_c; p.n_ctd = d->n_ctd;
  p.n_e = d->n_e; p.n_out = d->n_out; p.nsteps = d->nsteps; p.flags = d->flags;
  p.t0 = d->t0; p.dt = d->dt; p.atol = d->atol; p.rtol = d->rtol;
  p.max_step = d->max_step; p.first_step = d->first_step; p.min_step = d->min_step;
  p.probe_err = e->probe_err; p.stay_err = e->stay_err;
  // Kets: the weighted RMS norm averages over N components of which a handful carry the state, so the controller admits
  // local errors several times rtol on those, and over the many free (uncapped) steps of a ket solve the embedded pair ends
  // up further from the converged solution than the reference's variable-order Adams integrator does at the same
  // tolerances (five-qutrit problem of test_qasm.py:463-530: expectation traces 9e-5 vs 1.7e-5).  A factor 10 on the
  // internal tolerances (1.6x the steps of an N-vector integration) puts the ket path at or below the reference
  // integrator's error there; density matrices keep the caller's tolerances as they are.
  if (ket) { p.atol *= SEQ_KET_TOL_MARGIN; p.rtol *= SEQ_KET_TOL_MARGIN; }
  {      // experiments: thresholds from the environment, looked up per solve
    const char* s1 = getenv("SEQ_STAY_ERR");
    const char* s2 = getenv("SEQ_PROBE_ERR");
    if (s1 && atof(s1) > 0) p.stay_err = atof(s1);
    if (s2 && atof(s2) > 0) p.probe_err = atof(s2);
    if (p.probe_err > p.stay_err) p.probe_err = p.stay_err;
  }
  p.G0 = G0; p.G0_bstride = d->H0_batch_stride ? (long long)N2 : 0;
  p.Gk = Gk;
  p.tab = tabc; p.tab_rate = tabr;
  p.tab_bstride_c = d->coeff_batch_stride ? (long long)d->n_ch * d->n_t : 0;
  p.tab_bstride_r = d->rate_batch_stride ? (long long)d->n_ctd * d->n_t : 0;
  p.coeff_scale = d->coeff_scale;
  p.L = Lall;
  p.state0 = (const cplx*)d->state0; p.E = (const cplx*)d->E; p.out_idx = d->out_idx;
  p.expect = d->expect; p.final_state = (cplx*)d->final_state; p.states = (cplx*)d->states;
  p.status = d->status; p.stats = d->stats;
  p.ws = (cplx*)e->ws.ptr; p.ws_stride = (long long)ws_per;

  bool ket_ran = false;
  if (ket && ket_diag_wanted) {
    rc = launch_ket(e, d, p, G0, Gk, d->method == SEQ_METHOD_DIAG, &ket_ran);
    if (rc) return rc;
    if (ket_ran) e->last_method = SEQ_METHOD_DIAG;
    else if (d->method == SEQ_METHOD_DIAG) return fail(e, SEQ_ERR_UNSUPPORTED, "SEQ_METHOD_DIAG (kets): the operators are not sums of a few diagonals (N=%d)", N);
  }
  if (method != SEQ_METHOD_SPARSE && method != SEQ_METHOD_DIAG && !ket_ran) CUDA_TRY(e, cudaEventRecord(e->ev0, e->stream));
  if (method == SEQ_METHOD_SPARSE) {
    rc = launch_sparse(e, d, p, G0, Gk, spi);
    if (rc) return rc;
  } else if (method == SEQ_METHOD_DIAG) {
    rc = launch_diag(e, d, p, G0, Gk, spi, dpl, tail_ok ? &dpl_tail : nullptr, blk_ok ? &bpl : nullptr);
    if (rc) return rc;
  } else if (method == SEQ_METHOD_GENERIC && ket && ket_ran) {
    // (launched above)
  } else if (method == SEQ_METHOD_GENERIC) {
    size_t elems = ket ? (size_t)N : N2;
    int threads = (int)((elems + 31) / 32 * 32);
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    size_t smem = (34 + 64) * sizeof(double);
    if (ket) solve_generic_kernel<true><<<B, threads, smem, e->stream>>>(p);
    else solve_generic_kernel<false><<<B, threads, smem, e->stream>>>(p);
    e->launches++;
  } else if (method == SEQ_METHOD_DMMA) {
    rc = launch_dmma(p, G0, Gk, (double*)e->pack.ptr, e->stream, &e->launches);
    if (rc) return fail(e, rc, "DMMA kernel launch configuration failed for N=%d", N);
  } else if (method == SEQ_METHOD_DMMA_STREAM) {
    rc = launch_stream(p, G0, Gk, (double*)e->pack.ptr, e->stream, &e->launches);
    if (rc) return fail(e, rc, "streaming DMMA kernel launch configuration failed for N=%d", N);
  } else if (method == SEQ_METHOD_LARGE) {
    rc = solve_large(e, d, p, G0, Gk);
    if (rc) return rc;
  } else {
    rc = launch_small(p, G0, Gk, (double*)e->pack.ptr, e->sm_count, e->stream, &e->launches);
    if (rc) return fail(e, rc, "small-N kernel launch configuration failed for N=%d", N);
  }
  CUDA_TRY(e, cudaEventRecord(e->ev1, e->stream));
  e->timed = true;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

// end-to-end call on host buffers: stage in, solve, stage out (all on the engine's stream), synchronise if asked to
static int solve_host_one(seq_engine* e, const seq_solve_desc* d, bool sync) {
  const size_t N2 = (size_t)d->N * d->N, NN = state_elems(d), B = d->B;
  const size_t cz = sizeof(cplx);
  struct Seg { const void* src; size_t bytes; size_t off; };
  std::vector<Seg> in;
  size_t off = 0;
  auto add_in = [&](const void* src, size_t bytes) -> size_t {
    size_t o = off;
    if (src && bytes) { in.push_back({src, bytes, o}); off += (bytes + 255) / 256 * 256; }
    return o;
  };
  size_t oH0 = add_in(d->H0, (d->H0_batch_stride ? B : 1) * N2 * cz);
  size_t oHk = add_in(d->Hk, (size_t)d->n_ch * N2 * cz);
  size_t oco = add_in(d->coeff, (d->coeff_batch_stride ? B : 1) * (size_t)d->n_ch * d->n_t * sizeof(double));
  size_t ocs = add_in(d->coeff_scale, B * (size_t)d->n_ch * sizeof(double));
  size_t oLc = add_in(d->Lc, (size_t)d->n_c * N2 * cz);
  size_t oLt = add_in(d->Ltd, (size_t)d->n_ctd * N2 * cz);
  size_t ora = add_in(d->rate, (d->rate_batch_stride ? B : 1) * (size_t)d->n_ctd * d->n_t * sizeof(double));
  size_t os0 = add_in(d->state0, B * NN * cz);
  size_t oE = add_in(d->E, (size_t)d->n_e * N2 * cz);
  size_t ooi = add_in(d->out_idx, (size_t)d->n_out * sizeof(int32_t));
  int rc = reserve(e, e->stage_in, off + 256);
  if (rc) return rc;
  char* din = (char*)e->stage_in.ptr;
  for (const Seg& s : in) CUDA_TRY(e, cudaMemcpyAsync(din + s.off, s.src, s.bytes, cudaMemcpyHostToDevice, e->stream));

  const size_t esz = (d->flags & SEQ_FLAG_EXPECT_COMPLEX) ? cz : sizeof(double);
  size_t b_exp = B * (size_t)d->n_e * d->n_out * esz;
  size_t b_fin = B * NN * cz;
  size_t b_sts = (d->flags & SEQ_FLAG_STORE_STATES) ? B * (size_t)d->n_out * NN * cz : 0;
  size_t b_status = B * sizeof(int32_t), b_stats = B * 4 * sizeof(int32_t);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o_exp = 0, o_fin = o_exp + up(b_exp), o_sts = o_fin + up(b_fin), o_status = o_sts + up(b_sts), o_stats = o_status + up(b_status);
  rc = reserve(e, e->stage_out, o_stats + up(b_stats) + 256);
  if (rc) return rc;
  char* dout = (char*)e->stage_out.ptr;

  seq_solve_desc dd = *d;
  dd.flags &= ~SEQ_FLAG_HOST_POINTERS;
  dd.H0 = d->H0 ? din + oH0 : nullptr;
  dd.Hk = d->Hk && d->n_ch ? din + oHk : nullptr;
  dd.coeff = d->coeff && d->n_ch ? (const double*)(din + oco) : nullptr;
  dd.coeff_scale = d->coeff_scale && d->n_ch ? (const double*)(din + ocs) : nullptr;
  dd.Lc = d->Lc && d->n_c ? din + oLc : nullptr;
  dd.Ltd = d->Ltd && d->n_ctd ? din + oLt : nullptr;
  dd.rate = d->rate && d->n_ctd ? (const double*)(din + ora) : nullptr;
  dd.state0 = din + os0;
  dd.E = d->E && d->n_e ? din + oE : nullptr;
  dd.out_idx = (const int32_t*)(din + ooi);
  dd.expect = d->n_e ? dout + o_exp : nullptr;
  dd.final_state = dout + o_fin;
  dd.states = b_sts ? dout + o_sts : nullptr;
  dd.status = (int32_t*)(dout + o_status);
  dd.stats = (int32_t*)(dout + o_stats);
  rc = solve_device(e, &dd);
  if (rc) return rc;
  if (b_exp) CUDA_TRY(e, cudaMemcpyAsync(d->expect, dout + o_exp, b_exp, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaMemcpyAsync(d->final_state, dout + o_fin, b_fin, cudaMemcpyDeviceToHost, e->stream));
  if (b_sts) CUDA_TRY(e, cudaMemcpyAsync(d->states, dout + o_sts, b_sts, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaMemcpyAsync(d->status, dout + o_status, b_status, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(e, cudaMemcpyAsync(d->stats, dout + o_stats, b_stats, cudaMemcpyDeviceToHost, e->stream));
  if (sync) CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  return SEQ_OK;
}

// How many chunks a host-buffer solve is cut into.  One CTA per trajectory (N > 6): every chunk keeps at least two
// waves of CTAs; thread-per-trajectory (N <= 6): at least eight full CTAs per SM.  SEQ_HOST_CHUNKS overrides.
static int host_chunks(const seq_engine* e, const seq_solve_desc* d) {
  const char* env = getenv("SEQ_HOST_CHUNKS");
  long long n;
  if (env && atoi(env) > 0) n = atoi(env);
  else {
    if (pick_method(e, d) == SEQ_METHOD_LARGE) return 1;
    const long long per = d->N > 6 ? 3LL * e->sm_count : 8LL * 128 * e->sm_count;      // (N > 6: 2 chunks for the 1024-point C2 sweep - measured best: 19.4 k vs 19.15 k with 3, 19.0 k unchunked)
    n = d->B / per;
    if (n > 4) n = 4;
  }
  if (n > d->B) n = d->B;
  return n < 1 ? 1 : (int)n;
}

// Host-buffer solve: the batch is cut into chunks that alternate between two internal engines; chunk c+1's H2D
// copies and chunk c-1's D2H copies run while chunk c integrates (the per-chunk set-up - operator analysis, spline
// preparation - is repeated per chunk: microseconds).  Results do not depend on the chunking: trajectories are
// independent and every chunk runs the same kernel.
static int solve_host(seq_engine* e, const seq_solve_desc* d) {
  const int nchunk = host_chunks(e, d);
  e->chunks = 1;
  if (nchunk <= 1) return solve_host_one(e, d, true);
  for (int q = 0; q < 2; q++) {
    if (e->sub[q]) continue;
    CUDA_TRY(e, cudaStreamCreateWithFlags(&e->sub_stream[q], cudaStreamNonBlocking));
    int rc = seq_engine_create(e->device, e->sub_stream[q], &e->sub[q]);
    if (rc) return fail(e, rc, "could not create the internal engine of a chunked host solve");
    e->sub[q]->is_sub = true;
  }
  if (!e->ev_fork) CUDA_TRY(e, cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(e, cudaEventRecord(e->ev_fork, e->stream));          // everything queued on the caller's stream comes first
  for (int q = 0; q < 2; q++) CUDA_TRY(e, cudaStreamWaitEvent(e->sub_stream[q], e->ev_fork, 0));
  const size_t NN = state_elems(d), cz = sizeof(cplx);
  const size_t esz = (d->flags & SEQ_FLAG_EXPECT_COMPLEX) ? cz : sizeof(double);
  int launches = 0, rc = SEQ_OK;
  float ms = 0.f;
  long long b0 = 0;
  // CTA-per-trajectory kernels (N > 6): chunks of whole waves of CTA slots (two CTAs per SM for the block-local diagonal
  // kernel, one otherwise - 2 x SMs is a multiple of both), the rest goes with the last chunk, where the kernel's own
  // partial-wave handling sees it; equal chunks of 341 trajectories would each end in a 0.15 wave of their own
  const long long wave = d->N > 6 ? 2LL * e->sm_count : 0;
  const long long full_waves = wave ? d->B / wave : 0;
  for (int c = 0; c < nchunk && rc == SEQ_OK; c++) {
    long long nb = (d->B - b0 + (nchunk - c) - 1) / (nchunk - c);
    if (wave && full_waves >= nchunk) {
      const long long w_c = full_waves / nchunk + (c < full_waves % nchunk ? 1 : 0);
      nb = (c == nchunk - 1) ? d->B - b0 : w_c * wave;
    }
    seq_engine* s = e->sub[c & 1];
    if (c >= 2) {                       // the handle's counters of chunk c-2 are about to be overwritten
      launches += s->launches;
      float m1 = seq_engine_last_kernel_ms(s);
      if (m1 > 0) ms += m1;
    }
    seq_solve_desc dd = *d;
    dd.B = (int32_t)nb;
    if (d->H0 && d->H0_batch_stride) dd.H0 = (const char*)d->H0 + (size_t)b0 * d->H0_batch_stride * cz;
    if (d->coeff && d->coeff_batch_stride) dd.coeff = d->coeff + (size_t)b0 * d->coeff_batch_stride;
    if (d->coeff_scale) dd.coeff_scale = d->coeff_scale + (size_t)b0 * d->n_ch;
    if (d->rate && d->rate_batch_stride) dd.rate = d->rate + (size_t)b0 * d->rate_batch_stride;
    dd.state0 = (const char*)d->state0 + (size_t)b0 * NN * cz;
    if (d->expect) dd.expect = (char*)d->expect + (size_t)b0 * d->n_e * d->n_out * esz;
    dd.final_state = (char*)d->final_state + (size_t)b0 * NN * cz;
    if (d->states) dd.states = (char*)d->states + (size_t)b0 * d->n_out * NN * cz;
    dd.status = d->status + b0;
    dd.stats = d->stats + 4 * b0;
    s->launches = 0;
    s->timed = false;
    s->comm = nullptr;
    if (dd.n_c > 0 && dd.n_ctd > 0) {   // as in seq_engine_solve: room for the packed collapse operators
      const size_t N2 = (size_t)dd.N * dd.N, nG0 = dd.H0_batch_stride ? (size_t)dd.B : 1;
      rc = reserve(s, s->gbuf, (nG0 + dd.n_ch + dd.n_ctd + dd.n_c + dd.n_ctd) * N2 * sizeof(cplx));
    }
    if (rc == SEQ_OK) rc = solve_host_one(s, &dd, false);
    if (rc != SEQ_OK) e->err = s->err;
    b0 += nb;
  }
  for (int q = 0; q < 2; q++) {
    cudaError_t ce = cudaStreamSynchronize(e->sub_stream[q]);
    if (ce != cudaSuccess && rc == SEQ_OK) rc = fail(e, SEQ_ERR_CUDA, "chunked host solve failed: %s", cudaGetErrorString(ce));
  }
  if (rc != SEQ_OK) return rc;
  for (int q = 0; q < 2 && q < nchunk; q++) {
    launches += e->sub[q]->launches;
    float m1 = seq_engine_last_kernel_ms(e->sub[q]);
    if (m1 > 0) ms += m1;
  }
  e->launches = launches;
  e->last_method = e->sub[0]->last_method;
  e->last_fma_per_rhs = e->sub[0]->last_fma_per_rhs;
  e->last_fp64_step_extra = e->sub[0]->last_fp64_step_extra;
  char buf[64];
  snprintf(buf, sizeof(buf), " [host solve pipelined in %d chunks]", nchunk);
  e->variant = e->sub[0]->variant + buf;
  e->chunks = nchunk;
  e->chunk_ms = ms;
  return SEQ_OK;
}

// NVTX range around every C-ABI solve (visible in Nsight Systems / ncu --nvtx) and, when SEQ_PERF_LOG names a file, one
// JSON line per call appended to it: shape, method, kernel variant, launches, integration-kernel time, wall time of the call
// (per-call perf record).  Logging synchronises the engine's stream to read the kernel time.
struct SolveScope {
  seq_engine* e; const seq_solve_desc* d; const char* what; int* rc;
  std::chrono::steady_clock::time_point t0;
  SolveScope(seq_engine* e_, const seq_solve_desc* d_, const char* what_, int* rc_) : e(e_), d(d_), what(what_), rc(rc_), t0(std::chrono::steady_clock::now()) {
    nvtxRangePushA(what_);
  }
  ~SolveScope() {
    nvtxRangePop();
    const char* path = getenv("SEQ_PERF_LOG");
    if (!path || !*path || !e || !d || (e->is_sub)) return;
    const float kms = (*rc == SEQ_OK) ? seq_engine_last_kernel_ms(e) : -1.0f;      // (synchronises on the kernel's end event)
    const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (FILE* f = fopen(path, "a")) {
      fprintf(f, "{\"call\": \"%s\", \"rc\": %d, \"N\": %d, \"B\": %d, \"n_t\": %d, \"n_ch\": %d, \"n_c\": %d, \"n_ctd\": %d, \"n_e\": %d, \"n_out\": %d, "
                 "\"ket\": %d, \"host_pointers\": %d, \"atol\": %.3g, \"rtol\": %.3g, \"method\": %d, \"variant\": \"%s\", \"launches\": %d, "
                 "\"kernel_ms\": %.4f, \"call_wall_ms\": %.4f}\n",
              what, *rc, d->N, d->B, d->n_t, d->n_ch, d->n_c, d->n_ctd, d->n_e, d->n_out, (d->flags & SEQ_FLAG_KET) ? 1 : 0,
              (d->flags & SEQ_FLAG_HOST_POINTERS) ? 1 : 0, d->atol, d->rtol, e->last_method, e->variant.c_str(), e->launches, kms, wall);
      fclose(f);
    }
  }
};

static int seq_engine_solve_impl(seq_engine* e, const seq_solve_desc* d);

int seq_engine_solve(seq_engine* e, const seq_solve_desc* d) {
  if (!e) return SEQ_ERR_INVALID;
  int rc = SEQ_OK;
  SolveScope scope(e, d, "seq_engine_solve", &rc);
  rc = seq_engine_solve_impl(e, d);
  return rc;
}

static int seq_engine_solve_impl(seq_engine* e, const seq_solve_desc* d) {
  int rc = validate(e, d);
  if (rc) return rc;
  e->launches = 0;
  e->timed = false;
  e->chunks = 1;
  if (d->B == 0) return SEQ_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  {
    const bool has_tables = d->n_ch + d->n_ctd > 0;
    if (d->flags & SEQ_FLAG_HOST_POINTERS) {
      rc = validate_out_idx(e, d->out_idx, d->n_out, d->n_t, has_tables);
    } else {
      std::vector<int32_t> h((size_t)d->n_out);
      CUDA_TRY(e, cudaMemcpyAsync(h.data(), d->out_idx, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      rc = validate_out_idx(e, h.data(), d->n_out, d->n_t, has_tables);
    }
    if (rc) return rc;
  }
  // gbuf must also hold the packed collapse operators when both kinds are present
  if (d->n_c > 0 && d->n_ctd > 0) {
    size_t N2 = (size_t)d->N * d->N;
    size_t nG0 = d->H0_batch_stride ? (size_t)d->B : 1;
    rc = reserve(e, e->gbuf, (nG0 + d->n_ch + d->n_ctd + d->n_c + d->n_ctd) * N2 * sizeof(cplx));
    if (rc) return rc;
  }
  if (d->flags & SEQ_FLAG_HOST_POINTERS) return solve_host(e, d);
  return solve_device(e, d);
}

int seq_engine_solve_sharded(seq_engine* e, const seq_solve_desc* d, const seq_comm* comm) {
  if (!e) return SEQ_ERR_INVALID;
  if (!comm && e->nccl_comm && e->nccl_world > 1) comm = &e->nccl_hooks;      // the native communicator
  if (comm && comm->world > 1) {
    if (comm->rank < 0 || comm->rank >= comm->world || !comm->all_gather || !comm->all_reduce_sum)
      return fail(e, SEQ_ERR_INVALID, "bad communicator (rank %d of %d, hooks %p %p)", comm->rank, comm->world,
                  (void*)comm->all_gather, (void*)comm->all_reduce_sum);
    if (d && d->method != SEQ_METHOD_AUTO && d->method != SEQ_METHOD_LARGE)
      return fail(e, SEQ_ERR_UNSUPPORTED, "row-sharded solves run SEQ_METHOD_LARGE only");
  }
  e->comm = comm;
  seq_solve_desc dd;
  const seq_solve_desc* use = d;
  if (d && d->size == sizeof(seq_solve_desc) && comm && comm->world > 1) { dd = *d; dd.method = SEQ_METHOD_LARGE; use = &dd; }
  int rc = seq_engine_solve(e, use);
  e->comm = nullptr;
  return rc;
}

#ifdef SEQ_PROFILE_REGIONS
// profiling builds only (tools/region_prof.py): cycles per region of solve_dmma_kernel's step loop
int seq_engine_debug_regions(unsigned long long* out16, int reset) {
  if (out16 && cudaMemcpyFromSymbol(out16, seq::seq_prof, 16 * sizeof(unsigned long long)) != cudaSuccess) return SEQ_ERR_CUDA;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(seq::seq_prof, z, sizeof(z)) != cudaSuccess) return SEQ_ERR_CUDA;
  }
  return SEQ_OK;
}
#endif

int seq_engine_synth_gaussian(seq_engine* e, const seq_pulse_desc* descs, int64_t n_desc, double* coeff, int32_t B, int32_t n_ch,
                              int32_t n_t, int accumulate) {
  if (!e) return SEQ_ERR_INVALID;
  if (n_desc < 0 || (n_desc > 0 && !descs) || !coeff || B < 0 || n_ch < 1 || n_t < 1) return fail(e, SEQ_ERR_INVALID, "bad synth_gaussian arguments");
  for (int64_t q = 0; q < n_desc; q++)
    if (!(descs[q].sigma > 0) || !(descs[q].dt > 0) || !(descs[q].chop > 0))
      return fail(e, SEQ_ERR_INVALID, "pulse descriptor %lld: sigma, chop and dt must be positive", (long long)q);
  e->launches = 0;
  CUDA_TRY(e, cudaSetDevice(e->device));
  const size_t bytes = (size_t)B * n_ch * n_t * sizeof(double);
  if (!accumulate && bytes) CUDA_TRY(e, cudaMemsetAsync(coeff, 0, bytes, e->stream));
  if (n_desc == 0 || B == 0) return SEQ_OK;
  int rc = reserve(e, e->stage_in, (size_t)n_desc * sizeof(seq_pulse_desc) + 256);
  if (rc) return rc;
  CUDA_TRY(e, cudaMemcpyAsync(e->stage_in.ptr, descs, (size_t)n_desc * sizeof(seq_pulse_desc), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));    // `descs` is caller memory (possibly pageable)
  synth_gaussian_kernel<<<(unsigned)n_desc, 128, 0, e->stream>>>((const seq_pulse_desc*)e->stage_in.ptr, coeff, B, n_ch, n_t);
  e->launches++;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

int seq_engine_apply_unitary(seq_engine* e, const void* U, void* states, int32_t N, int32_t B, int is_ket) {
  if (!e) return SEQ_ERR_INVALID;
  if (!U || !states || N < 1 || B < 0) return fail(e, SEQ_ERR_INVALID, "bad apply_unitary arguments");
  e->launches = 0;
  if (B == 0) return SEQ_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  size_t per = is_ket ? (size_t)N : (size_t)N * N;
  int rc = reserve(e, e->ws, per * B * sizeof(cplx));
  if (rc) return rc;
  int threads = (int)((per + 31) / 32 * 32);
  if (threads > 256) threads = 256;
  apply_unitary_kernel<<<B, threads, 0, e->stream>>>((const cplx*)U, (cplx*)states, (cplx*)e->ws.ptr, N, is_ket);
  e->launches++;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

int seq_engine_expect(seq_engine* e, const void* E, const void* states, void* out, int32_t N, int32_t B, int32_t n_e,
                      int is_ket) {
  if (!e) return SEQ_ERR_INVALID;
  if (!E || !states || !out || N < 1 || B < 0 || n_e < 0) return fail(e, SEQ_ERR_INVALID, "bad expect arguments");
  e->launches = 0;
  if (B == 0 || n_e == 0) return SEQ_OK;
  CUDA_TRY(e, cudaSetDevice(e->device));
  size_t per = is_ket ? (size_t)N : (size_t)N * N;
  int threads = (int)((per + 31) / 32 * 32);
  if (threads > 256) threads = 256;
  dim3 grid(B, n_e);
  expect_kernel<<<grid, threads, 0, e->stream>>>((const cplx*)E, (const cplx*)states, (cplx*)out, N, n_e, is_ket);
  e->launches++;
  CUDA_TRY(e, cudaGetLastError());
  return SEQ_OK;
}

}  // extern "C"
