// Instantiations of the streaming kernel for one harmonic count (compile with -DVCB_INST_H=<0..5>).
#include "vcb_stream.cuh"

#ifndef VCB_INST_H
#error "compile with -DVCB_INST_H=<number of gene harmonics>"
#endif

namespace vcb {

template <int H, bool VELO, bool GRAD, bool LGI, int NP>
static cudaError_t launch_t(const StreamParams& sp, dim3 grid, int nthr, int smem, cudaStream_t st) {
  auto kfn = vcb_stream_kernel<H, VELO, GRAD, LGI, NP>;
  static int configured_smem = 0;  // per instantiation; raising the opt-in limit is idempotent
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_smem = smem;
  }
  kfn<<<grid, nthr, smem, st>>>(sp);
  return cudaGetLastError();
}

template <int H, int NP>
static cudaError_t launch_np(bool velo, bool grad, bool lgi, const StreamParams& sp, dim3 grid, int nthr, int smem,
                             cudaStream_t st) {
  if (velo) {
    if (grad) return lgi ? launch_t<H, true, true, true, NP>(sp, grid, nthr, smem, st)
                         : launch_t<H, true, true, false, NP>(sp, grid, nthr, smem, st);
    return lgi ? launch_t<H, true, false, true, NP>(sp, grid, nthr, smem, st)
               : launch_t<H, true, false, false, NP>(sp, grid, nthr, smem, st);
  }
  if (grad) return lgi ? launch_t<H, false, true, true, NP>(sp, grid, nthr, smem, st)
                       : launch_t<H, false, true, false, NP>(sp, grid, nthr, smem, st);
  return lgi ? launch_t<H, false, false, true, NP>(sp, grid, nthr, smem, st)
             : launch_t<H, false, false, false, NP>(sp, grid, nthr, smem, st);
}

#define VCB_CAT_(a, b) a##b
#define VCB_CAT(a, b) VCB_CAT_(a, b)

cudaError_t VCB_CAT(vcb_launch_stream_h, VCB_INST_H)(bool velo, bool grad, bool lgi, int np, const StreamParams& sp,
                                                     dim3 grid, int nthr, int smem, cudaStream_t st) {
  if (np == 2) return launch_np<VCB_INST_H, 2>(velo, grad, lgi, sp, grid, nthr, smem, st);
  return launch_np<VCB_INST_H, 1>(velo, grad, lgi, sp, grid, nthr, smem, st);
}

}  // namespace vcb
