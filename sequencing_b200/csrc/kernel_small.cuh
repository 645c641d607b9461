// placeholder until the small-N kernel lands
#pragma once
#include "common.cuh"
namespace seq {
static inline bool small_supported(int N, int n_g, int n_l) { (void)N; (void)n_g; (void)n_l; return false; }
static inline int launch_small(const SolveParams& p, int sms, cudaStream_t s) { (void)p; (void)sms; (void)s; return SEQ_ERR_UNSUPPORTED; }
}
