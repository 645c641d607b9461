// Device-resident step control for the large path's structured right-hand side (one GPU): the embedded-RK stepper of
// seq_engine.cu::solve_large without its per-step D2H copy of the error norm and stream synchronisation.  The
// controller (the shared StepCtl) runs in a one-warp kernel and leaves everything a step needs - step size, pair,
// first stage, spline values of every stage time, which buffers hold y / k1, whether an output is due - in a control
// block in device memory; every stage kernel reads its parameters from that block and returns at once when it has
// nothing to do (a stage the pair does not have, a finished solve).  The host enqueues ATTEMPTS (controller, output,
// up to seven stages, error norm) ahead of the GPU and looks at the block once per batch of attempts.
#pragma once
#include "kernel_large.cuh"
#include "kernel_large_diag.cuh"

namespace seq {

#define LDEV_MAXG 64

struct LargeDev {
  StepCtl ctl;
  double htry, target;
  int o, nst, landing, capped;
  int active;              // an attempt is being evaluated
  int S, st0;
  int emit, emit_o;        // the committed state is output `emit_o` (emit != 0)
  int ycur;                // ybuf[ycur] holds y, ybuf[ycur ^ 1] receives the new state
  int kmap[7];             // buffer of stage derivative j
  int attempts;
  double err2;             // sum of the weighted squared errors of the attempt (large_err accumulates into it)
  double cval[7 * LDEV_MAXG];
};

struct LargeDevBufs {
  double* ybuf[2];
  double* kbuf[7];
  double* Yfull;
  long long mplane, yplane;      // doubles per plane of a row-block vector / of the full stage buffer
};

// Programmatic dependent launch (the attempt's kernels are launched with programmaticStreamSerialization): a kernel's CTAs
// become resident while the previous kernel drains; nothing written by an earlier kernel is read before this wait.  The
// dependents are released right after it, so a kernel two launches back has always completed.  (No-ops without the attribute.)
__device__ __forceinline__ void lh_grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// launch with programmatic stream serialization (see lh_grid_dependency_wait)
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(bool pdl, void (*k)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}

// One warp.  `first`: nothing has been evaluated yet (initial call).
__global__ void ldev_ctl_kernel(const __grid_constant__ SolveParams p, int b, LargeDev* __restrict__ dev, int first) {
  const int lane = threadIdx.x;
  lh_grid_dependency_wait();
  StepCtl ctl = dev->ctl;
  double c_htry = dev->htry, c_target = dev->target;
  int c_o = dev->o, c_nst = dev->nst;
  bool c_landing = dev->landing != 0, c_capped = dev->capped != 0;
  const int was_active = dev->active;
  __syncwarp();
  if (!first && !was_active) { if (lane == 0) dev->emit = 0; return; }      // finished earlier: nothing to do
  bool accepted = false;
  int S_prev = dev->S, st0_prev = dev->st0;
  if (!first) {
    ctl.nrhs += S_prev - st0_prev;
    ctl.have_k1 = true;
    const double err = sqrt(dev->err2 / ((double)p.N * (double)p.N));
    accepted = ctl_update(ctl, p, err, c_htry, c_landing, c_capped, c_target, c_nst);
  }
  int emit = 0, emit_o = 0, cmd = 1;
  const int emit_lo = c_o;
  while (ctl.status == SEQ_STATUS_OK && c_o < p.n_out) {
    c_target = p.t0 + p.dt * (double)p.out_idx[c_o];
    if (c_o > 0 && ctl_next(ctl, p, c_target, c_htry, c_landing, c_capped)) { cmd = 0; break; }
    c_o++;
    c_nst = 0;
  }
  if (ctl.status == SEQ_STATUS_OK && c_o > emit_lo) { emit = 1; emit_o = c_o - 1; }     // (output samples are distinct: at most one per step)
  const int pair = ctl.pair, S = pair ? 5 : 7, st0 = ctl.have_k1 ? 1 : 0;
  if (cmd == 0) {
    const double2* tabc = p.tab + (long long)b * p.tab_bstride_c;
    const double2* tabr = p.tab_rate + (long long)b * p.tab_bstride_r;
    for (int q = lane; q < S * p.n_g; q += 32) {
      const int st = q / p.n_g, m = q - st * p.n_g;
      const double ts = ctl.t + RK_C[pair][st] * c_htry;
      double v;
      if (m < p.n_ch) {
        v = spline_eval(tabc + (long long)m * p.n_t, p.n_t, p.t0, p.dt, ts);
        if (p.coeff_scale) v *= p.coeff_scale[(long long)b * p.n_ch + m];
      } else {
        v = spline_eval(tabr + (long long)(m - p.n_ch) * p.n_t, p.n_t, p.t0, p.dt, ts);
      }
      dev->cval[st * LDEV_MAXG + m] = v;
    }
  }
  if (lane == 0) {
    if (accepted) {      // commit: the new state and the derivative of the last stage become y and k1
      dev->ycur ^= 1;
      const int t = dev->kmap[0]; dev->kmap[0] = dev->kmap[S_prev - 1]; dev->kmap[S_prev - 1] = t;
    }
    dev->ctl = ctl;
    dev->htry = c_htry; dev->target = c_target;
    dev->o = c_o; dev->nst = c_nst; dev->landing = c_landing ? 1 : 0; dev->capped = c_capped ? 1 : 0;
    dev->active = (cmd == 0) ? 1 : 0;
    dev->S = S; dev->st0 = st0;
    dev->emit = emit; dev->emit_o = emit_o;
    dev->err2 = 0.0;
    dev->attempts += (cmd == 0) ? 1 : 0;
  }
}

// stage input of stage `st`: y + h sum_j a_j k_j -> the full stage buffer (and, last stage, the new-state buffer)
__global__ void ldev_combine_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeDevBufs bf, int st) {
  if (!dev->active || st < dev->st0 || st >= dev->S || st == 0) return;      // (stage 0 reads y itself: the stage buffer already holds it)
  const int pair = dev->ctl.pair, S = dev->S;
  const double h = dev->htry;
  const double* y = bf.ybuf[dev->ycur];
  double* keep = (st == S - 1) ? bf.ybuf[dev->ycur ^ 1] : nullptr;
  const double* k[7];
  double a[7];
  int nk = 0;
  for (int j = 0; j < st; j++) {
    const double aj = RK_A[pair][st][j];
    k[j] = bf.kbuf[dev->kmap[j]]; a[j] = aj;
    nk = j + 1;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * bf.mplane; i += (long long)gridDim.x * blockDim.x) {
    const int pl = i >= bf.mplane;
    const long long r = i - pl * bf.mplane;
    double acc = 0.0;
    for (int j = 0; j < nk; j++) acc = fma(a[j], k[j][i], acc);
    const double v = fma(h, acc, y[i]);
    bf.Yfull[pl * bf.yplane + r] = v;
    if (keep) keep[i] = v;
  }
}

__global__ void ldev_gt_kernel(const LargeDev* __restrict__ dev, int st, const cplx* __restrict__ g0, const cplx* __restrict__ gk, int n_g, int n,
                               cplx* __restrict__ gt) {
  if (!dev->active || st < dev->st0 || st >= dev->S) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double* cval = dev->cval + st * LDEV_MAXG;
  cplx g = g0[q];
  for (int m = 0; m < n_g; m++) {
    const cplx a = gk[(long long)m * n + q];
    const double v = cval[m];
    g.x = fma(v, a.x, g.x);
    g.y = fma(v, a.y, g.y);
  }
  gt[q] = g;
}

__global__ void __launch_bounds__(256) ldev_rhs_kernel(const LargeDev* __restrict__ dev, const __grid_constant__ LargeDevBufs bf, int st,
                                                       const __grid_constant__ LargeDiagArgs a0) {
  if (!dev->active || st < dev->st0 || st >= dev->S) return;
  large_diag_rhs_body(a0, bf.kbuf[dev->kmap[st]], dev->cval + st * LDEV_MAXG);
}

__global__ void ldev_err_kernel(LargeDev* __restrict__ dev, const __grid_constant__ LargeDevBufs bf, double atol, double rtol, int m, int NP, int N) {
  __shared__ double red[34];
  if (!dev->active) return;
  const int pair = dev->ctl.pair, S = dev->S;
  const double h = dev->htry;
  const double* y = bf.ybuf[dev->ycur];
  const double* yn = bf.ybuf[dev->ycur ^ 1];
  const double* k[7];
  double ec[7];
  for (int j = 0; j < S; j++) { k[j] = bf.kbuf[dev->kmap[j]]; ec[j] = RK_E[pair][j]; }
  const long long n_plane = (long long)m * NP;
  double es = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_plane; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / NP;
    const int col = (int)(i - r * NP);
    if (r >= N || col >= N) continue;
    double er = 0.0, ei = 0.0;
    for (int j = 0; j < S; j++) { er = fma(ec[j], k[j][i], er); ei = fma(ec[j], k[j][n_plane + i], ei); }
    er *= h; ei *= h;
    const double y2 = y[i] * y[i] + y[n_plane + i] * y[n_plane + i];
    const double n2 = yn[i] * yn[i] + yn[n_plane + i] * yn[n_plane + i];
    const double s = atol + rtol * sqrt(fmax(y2, n2));
    es += (er * er + ei * ei) / (s * s);
  }
  es = block_sum(es, red);
  if (threadIdx.x == 0) atomicAdd(&dev->err2, es);
}

// Tr(E_e rho) of the committed state from COO lists of the expectation operators (E_ij at (i, j): sum E_ij rho_ji)
__device__ __forceinline__ void ldev_expect_body(const LargeDev* __restrict__ dev, const int e, const int* __restrict__ e_off, const unsigned* __restrict__ e_ij,
                                                 const cplx* __restrict__ e_val, const double* __restrict__ Y, long long y_plane, int NP, void* __restrict__ expect,
                                                 long long base, int n_out, int complex_out, double* red) {
  if (!dev->emit) return;
  cplx acc = cmake(0, 0);
  for (int q = e_off[e] + threadIdx.x; q < e_off[e + 1]; q += blockDim.x) {
    const unsigned code = e_ij[q];
    const long long i = code & 0xffffu, j = code >> 16;
    const cplx r = cmake(Y[j * NP + i], Y[y_plane + j * NP + i]);      // rho_ji
    cfma(acc, e_val[q], r);
  }
  acc = block_sum_c(acc, red);
  if (threadIdx.x == 0) {
    const long long off = base + (long long)e * n_out + dev->emit_o;
    if (complex_out) ((cplx*)expect)[off] = acc; else ((double*)expect)[off] = acc.x;
  }
}
__global__ void ldev_expect_kernel(const LargeDev* __restrict__ dev, const int* __restrict__ e_off, const unsigned* __restrict__ e_ij,
                                   const cplx* __restrict__ e_val, const double* __restrict__ Y, long long y_plane, int NP, void* __restrict__ expect,
                                   long long base, int n_out, int complex_out) {
  __shared__ double red[34];
  ldev_expect_body(dev, blockIdx.x, e_off, e_ij, e_val, Y, y_plane, NP, expect, base, n_out, complex_out, red);
}

__global__ void ldev_store_kernel(const LargeDev* __restrict__ dev, const double* __restrict__ src, cplx* __restrict__ states, int N, int NP, long long rows_pad) {
  if (!dev->emit) return;
  cplx* dst = states + (long long)dev->emit_o * N * N;
  const long long plane = rows_pad * NP;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)N * N; idx += (long long)gridDim.x * blockDim.x) {
    const long long i = idx / N;
    const int j = (int)(idx - i * N);
    dst[idx] = cmake(src[i * NP + j], src[plane + i * NP + j]);
  }
}

}  // namespace seq
