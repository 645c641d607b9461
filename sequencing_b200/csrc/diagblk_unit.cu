// Translation unit of the block-local diagonal kernel (kernel_diagblk.cuh): its template instantiations compile in
// parallel with seq_engine.cu.  The headers are included inside a renamed namespace so that the non-template kernels
// and __constant__ tables they define do not collide with seq_engine.cu's copies at link time; the structs that cross
// the boundary (SolveParams, BlkPlan) are the same definitions on both sides.
#define seq seq_blk_unit
#ifndef SEQ_PROFILE_REGIONS
#define PROF_DECL
#define PROF(r)
#else
#error "region profiling builds compile everything in one translation unit (no -DSEQ_BLK_EXTERN)"
#endif
#include "kernel_diagblk.cuh"

extern "C" int seq_blk_launch_ext(const void* p, const void* bp, void* stream) {
  return seq_blk_unit::launch_diagblk_kernel(*(const seq_blk_unit::SolveParams*)p, *(const seq_blk_unit::BlkPlan*)bp, (cudaStream_t)stream);
}
