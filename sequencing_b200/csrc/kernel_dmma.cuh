// placeholder until the DMMA kernel lands
#pragma once
#include "common.cuh"
namespace seq {
static inline bool dmma_supported(int N, int max_smem) { (void)N; (void)max_smem; return false; }
static inline size_t dmma_ws_elems(int N) { (void)N; return 0; }
static inline int launch_dmma(const SolveParams& p, int sms, int max_smem, cudaStream_t s) { (void)p; (void)sms; (void)max_smem; (void)s; return SEQ_ERR_UNSUPPORTED; }
}
