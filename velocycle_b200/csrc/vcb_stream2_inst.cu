// Instantiations of the round-2 streaming kernel for one harmonic count (compile with -DVCB_INST_H=<0..5>).
#include "vcb_stream2.cuh"

#ifndef VCB_INST_H
#error "compile with -DVCB_INST_H=<number of gene harmonics>"
#endif

namespace vcb {

template <int H, bool VELO, bool GRAD>
static cudaError_t launch2_t(const s2::Params& sp, dim3 grid, int nthr, int smem, cudaStream_t st) {
  auto kfn = s2::vcb_stream2_kernel<H, VELO, GRAD>;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  // idempotent, per device
  if (e != cudaSuccess) return e;
  kfn<<<grid, nthr, smem, st>>>(sp);
  return cudaGetLastError();
}

#define VCB_CAT_(a, b) a##b
#define VCB_CAT(a, b) VCB_CAT_(a, b)

cudaError_t VCB_CAT(vcb_launch_stream2_h, VCB_INST_H)(bool velo, bool grad, const s2::Params& sp, dim3 grid, int nthr, int smem,
                                                      cudaStream_t st) {
  if (velo) return grad ? launch2_t<VCB_INST_H, true, true>(sp, grid, nthr, smem, st)
                        : launch2_t<VCB_INST_H, true, false>(sp, grid, nthr, smem, st);
  return grad ? launch2_t<VCB_INST_H, false, true>(sp, grid, nthr, smem, st)
              : launch2_t<VCB_INST_H, false, false>(sp, grid, nthr, smem, st);
}

}  // namespace vcb
